#!/bin/bash
# round 2, single GPU: tensor-map staging after the alignment fix, hooks, slab tests on one GPU, V = 4 chunk sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hooks.py tests/test_gpu_multi.py -m gpu -q --timeout 900 -k "staging or stream_kernel or hooks or step_fun or convolver or multi or group or devices or slab" > gpurun_out/r03_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r03_pytest_gpu_b.log
{
for tma in 0 2; do
python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 96 --tma $tma
python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 48 --mask random --tma $tma
done
python tools/tune.py --n 16384 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0,256,512,1024,2048,4096 --vecs 2,4 --sweeps 48 --mask random
} > gpurun_out/r03_tune_tma2d_b.log 2>&1
echo "tune rc=$?"; grep GLUPS gpurun_out/r03_tune_tma2d_b.log; grep -i error gpurun_out/r03_tune_tma2d_b.log | head -3
