#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.log 2>&1; echo "bench$N rc=$?"
grep metric gpurun_out/bench$N.log | cut -c1-300
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check$N.log 2>&1; echo "dist_check rc=$?"
tail -6 gpurun_out/dist_check$N.log
timeout 1200 $TR tests/scripts/c5_dist.py --size 32768 --dtypes f32,f64 > gpurun_out/c5_dist${N}_32768.log 2>&1; echo "c5 rc=$?"
grep config gpurun_out/c5_dist${N}_32768.log | cut -c1-560
timeout 600 $TR tests/scripts/c5_dist.py --size 32768 --dtypes f32 --exchange harmonic=36,ccst=24 > gpurun_out/c5_dist${N}_32768_E2.log 2>&1; echo "c5 E2 rc=$?"
grep config gpurun_out/c5_dist${N}_32768_E2.log | cut -c1-560
