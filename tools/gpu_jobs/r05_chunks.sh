#!/bin/bash
# round 2 (session 5): chunk height on many-wave grids (the picker's ch/8192 penalty), CCST fp64 K=3 and harmonic K=6
mkdir -p gpurun_out
{
HMI_DEBUG=1 python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0,256,512,768,1024,2048,4096 --vecs 2 --sweeps 24 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --n 32768 --methods harmonic --dtypes f64 --kernels stream --tbs 6 --chunks 0,512,1024,2048 --vecs 2 --sweeps 24 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --n 16384 --methods ccst --dtypes f64,f32 --kernels stream --tbs 3 --chunks 0,256,512,1024,2048 --vecs 2 --sweeps 48 --mask random 2>&1 | grep -v "norm=1" | sort -u
} > gpurun_out/r05_tune_chunk_height_large_grids.log 2>&1
cat gpurun_out/r05_tune_chunk_height_large_grids.log
