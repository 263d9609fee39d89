#!/bin/bash
# round 2, single GPU: whole GPU suite (tensor-map staging fixed, device pyramid, hooks), staging A/B, short bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r03_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r03_pytest_gpu_c.log
{
for tma in 0 2; do
python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 96 --tma $tma
python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 48 --mask random --tma $tma
done
} > gpurun_out/r03_tune_tma2d_c.log 2>&1
echo "tune rc=$?"; grep GLUPS gpurun_out/r03_tune_tma2d_c.log; grep -i error gpurun_out/r03_tune_tma2d_c.log | head -3
