#!/bin/bash
# round 2 (session 4): first run of the warp-specialised pipeline kernel (tma = 3): parity tests, then a tune sweep
mkdir -p gpurun_out; nvidia-smi > /dev/null 2>&1; sleep 2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "stream_kernel_matches_oracle or staging or shapes_and_chunking" > gpurun_out/r04_pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r04_pytest_pipe.log
{
for n in 4096 16384; do
m=tracks; [ $n = 16384 ] && m=random
timeout 300 python tools/tune.py --n $n --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 48 --mask $m --tma 0
timeout 300 python tools/tune.py --n $n --methods ccst --dtypes f64 --kernels stream --tbs 2,3,4 --chunks 0 --vecs 4,8 --sweeps 48 --mask $m --tma 3
timeout 300 python tools/tune.py --n $n --methods harmonic --dtypes f64 --kernels stream --tbs 6 --chunks 0 --vecs 2 --sweeps 48 --mask $m --tma 0
timeout 300 python tools/tune.py --n $n --methods harmonic --dtypes f64 --kernels stream --tbs 4,6,8 --chunks 0 --vecs 4,8 --sweeps 48 --mask $m --tma 3
timeout 300 python tools/tune.py --n $n --methods ccst,harmonic --dtypes f32 --kernels stream --tbs 3,4,6 --chunks 0 --vecs 2 --sweeps 48 --mask $m --tma 0
timeout 300 python tools/tune.py --n $n --methods ccst,harmonic --dtypes f32 --kernels stream --tbs 3,4,6 --chunks 0 --vecs 2 --sweeps 48 --mask $m --tma 3
done
timeout 300 python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 24 --mask random --tma 0
timeout 300 python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3,4 --chunks 0 --vecs 4,8 --sweeps 24 --mask random --tma 3
} > gpurun_out/r04_tune_pipe.log 2>&1
echo "tune rc=$?"; cat gpurun_out/r04_tune_pipe.log
