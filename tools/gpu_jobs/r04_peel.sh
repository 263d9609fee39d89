#!/bin/bash
# round 2 (session 4): peeled steady-state row loop of the streaming kernel: parity tests, tune at three sizes, long-run bench A/B (vec 2 / 4)
mkdir -p gpurun_out; nvidia-smi > /dev/null 2>&1; sleep 2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "stream_kernel or staging or shapes or two_slabs or mid_sized or full_size_properties_4096 or converged" > gpurun_out/r04_pytest_peel.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r04_pytest_peel.log
{
timeout 300 python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 96
timeout 300 python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 48 --mask random
timeout 300 python tools/tune.py --n 32768 --methods ccst,harmonic --dtypes f64 --kernels stream --tbs 3,6 --chunks 0 --vecs 2,4 --sweeps 24 --mask random
} > gpurun_out/r04_tune_peel.log 2>&1
echo "tune rc=$?"; cat gpurun_out/r04_tune_peel.log
for v in 0 4; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --vec $v > gpurun_out/r04_bench_peel_vec$v.json 2> gpurun_out/r04_bench_peel_vec$v.err; echo "bench vec=$v rc=$?"; cut -c1-420 gpurun_out/r04_bench_peel_vec$v.json
done
