#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for n in 4096 8192 16384; do
timeout 600 python tools/tune.py --n $n --methods tv --dtypes f32 --kernels stream --tbs 2,3 --chunks=0 --vecs 2 >> gpurun_out/tune_tv_k3.log 2>&1
done
cat gpurun_out/tune_tv_k3.log
