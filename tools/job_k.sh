#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 300 python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks=0 --vecs 2 > gpurun_out/tune_triple.log 2>&1
timeout 300 python tools/tune.py --n 4096 --methods tv,amle --dtypes f64,f32 --kernels stream --tbs 2 --chunks=0 --vecs 2 >> gpurun_out/tune_triple.log 2>&1
timeout 300 python tools/tune.py --n 16384 --methods ccst,harmonic,tv,amle --dtypes f64 --kernels stream --tbs 2,3,6 --chunks=0 --vecs 2 >> gpurun_out/tune_triple.log 2>&1
cat gpurun_out/tune_triple.log
timeout 120 python tests/scripts/small_latency.py > gpurun_out/small_latency.log 2>&1; cat gpurun_out/small_latency.log
