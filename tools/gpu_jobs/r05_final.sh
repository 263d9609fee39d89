#!/bin/bash
# round 2 (session 5): final regression — every GPU test, smoke, both bench arms; cost of a band launch by chunk height
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900) > gpurun_out/r05_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r05_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 4 --warmup 2 > gpurun_out/r05_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/r05_bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r05_bench_1gpu_final.json 2> gpurun_out/r05_bench_1gpu_final.err; echo "bench rc=$?"; cat gpurun_out/r05_bench_1gpu_final.json
HMI_DEBUG=1 timeout 300 python tools/band_launch.py 2>&1 | grep -v "norm=1" | sort -u > gpurun_out/r05_tune_band_launch.log; cat gpurun_out/r05_tune_band_launch.log
{
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods harmonic --dtypes f32 --kernels stream --tbs 8 --chunks 0,683,456,342,228,171 --vecs 2 --sweeps 96 --mask random
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods ccst --dtypes f32 --kernels stream --tbs 4 --chunks 0,683,456,342,228,171 --vecs 2 --sweeps 96 --mask random
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0,512,410,342,256,171 --vecs 2 --sweeps 96 --mask random
} 2>&1 | grep -v "norm=1" | sort -u > gpurun_out/r05_tune_slab_chunks.log; cat gpurun_out/r05_tune_slab_chunks.log
