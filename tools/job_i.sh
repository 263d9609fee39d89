#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for n in 4096 8192 16384; do
HMI_DEBUG=1 timeout 600 python tools/tune.py --n $n --methods tv,amle --dtypes f32,f64 --kernels stream --tbs 2,3 --chunks=0 --vecs 2 >> gpurun_out/tune_nl_k2b.log 2>> gpurun_out/tune_nl_k2b.err
done
cat gpurun_out/tune_nl_k2b.log
sort -u gpurun_out/tune_nl_k2b.err | grep "norm=0" | head -20
timeout 600 python tools/mg_bench.py --n 4096 --levels 4 --dtype f32 > gpurun_out/mg_bench.log 2>&1
cat gpurun_out/mg_bench.log
