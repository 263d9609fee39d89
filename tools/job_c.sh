#!/bin/bash
# scratch 2-GPU job: slab parity, C5 tool smoke at 8192, bench at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "dist_check rc=$?"
tail -6 gpurun_out/dist_check2.log
timeout 600 $TR tests/scripts/c5_dist.py --n 8192 --dtypes f64,f32 > gpurun_out/c5_dist2_8192.log 2>&1; echo "c5 rc=$?"
grep config gpurun_out/c5_dist2_8192.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 tests/scripts/c5_dist.py --n 8192 --dtypes f64,f32 > gpurun_out/c5_dist1_8192.log 2>&1; echo "c5(1) rc=$?"
grep config gpurun_out/c5_dist1_8192.log | cut -c1-700
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.log 2>&1; echo "bench2 rc=$?"
grep metric gpurun_out/bench2.log | cut -c1-900
