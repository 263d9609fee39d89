#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "dist_check rc=$?"
tail -6 gpurun_out/dist_check2.log
timeout 300 python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks=0 --vecs 2 > gpurun_out/tune_sym.log 2>&1
cat gpurun_out/tune_sym.log
