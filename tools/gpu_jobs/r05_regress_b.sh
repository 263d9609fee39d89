#!/bin/bash
# round 2 (session 5): regression after the fp32 temporal-block defaults (CCST 4, harmonic 8), fp64 depth check on 32768^2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r05_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r05_pytest_gpu_b.log
{
python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3,4 --chunks 0 --vecs 2 --sweeps 24 --mask random
python tools/tune.py --n 32768 --methods harmonic --dtypes f64 --kernels stream --tbs 6,8 --chunks 0 --vecs 2,4 --sweeps 48 --mask random
python tools/tune.py --n 4096 --methods ccst --dtypes f64 --kernels stream --tbs 3,4 --chunks 0 --vecs 2 --sweeps 96
} > gpurun_out/r05_tune_f64_temporal_block.log 2>&1
cat gpurun_out/r05_tune_f64_temporal_block.log
