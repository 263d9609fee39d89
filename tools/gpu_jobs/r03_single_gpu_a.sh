#!/bin/bash
# round 2, single GPU: full GPU test suite after the refactor, row-staging A/B (cp.async / 1-D bulk / 2-D TMA boxes), bench N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r03_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r03_pytest_gpu.log
{
for tma in 0 1 2; do
python tools/tune.py --n 4096 --methods ccst --dtypes f64,f32 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96 --tma $tma
python tools/tune.py --n 4096 --methods harmonic --dtypes f64,f32 --kernels stream --tbs 6 --chunks 0 --vecs 2 --sweeps 96 --tma $tma
python tools/tune.py --n 16384 --methods ccst --dtypes f64,f32 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 48 --mask random --tma $tma
python tools/tune.py --n 16384 --methods harmonic --dtypes f64,f32 --kernels stream --tbs 6 --chunks 0 --vecs 2 --sweeps 48 --mask random --tma $tma
done
} > gpurun_out/r03_tune_tma2d.log 2>&1
echo "tune rc=$?"; cat gpurun_out/r03_tune_tma2d.log
timeout 900 python bench.py --steps 3 --warmup 3 --write-hashes > gpurun_out/r03_bench_1gpu.json 2> gpurun_out/r03_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r03_bench_1gpu.err; cat gpurun_out/r03_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03_bench_ref.json 2>&1; echo "ref rc=$?"; cat gpurun_out/r03_bench_ref.json
