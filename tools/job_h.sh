#!/bin/bash
mkdir -p gpurun_out
for n in 4096 8192; do
HMI_DEBUG=1 timeout 600 python tools/tune.py --n $n --methods tv --dtypes f32,f64 --kernels stream --tbs 2 --chunks=0,-1,48,64,96,128,192,256,384 --vecs 2 >> gpurun_out/tune_tv_chunks.log 2>> gpurun_out/tune_tv_chunks.err
done
cat gpurun_out/tune_tv_chunks.log
sort -u gpurun_out/tune_tv_chunks.err | grep "norm=0" | grep "edge=[1-9]"
