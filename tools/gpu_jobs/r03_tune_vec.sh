#!/bin/bash
# round 2: fp64 cells per lane (vec 2 vs 4) with the current renaming/skew loop, at three sizes
mkdir -p gpurun_out
{
python tools/tune.py --n 4096 --methods ccst --dtypes f64 --kernels stream --tbs 2,3,4 --chunks 0 --vecs 2,4 --sweeps 96
python tools/tune.py --n 4096 --methods harmonic --dtypes f64 --kernels stream --tbs 4,6,7 --chunks 0 --vecs 2,4 --sweeps 96
python tools/tune.py --n 16384 --methods ccst --dtypes f64 --kernels stream --tbs 3,4 --chunks 0 --vecs 2,4 --sweeps 48 --mask random
python tools/tune.py --n 16384 --methods harmonic --dtypes f64 --kernels stream --tbs 6,7 --chunks 0 --vecs 2,4 --sweeps 48 --mask random
python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2,4 --sweeps 24 --mask random
} > gpurun_out/r03_tune_vec.log 2>&1
echo rc=$?; tail -40 gpurun_out/r03_tune_vec.log
