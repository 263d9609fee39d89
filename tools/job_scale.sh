#!/bin/bash
# scratch N-GPU job: bench at N, slab parity, C5 (32768^2) strong-scaling point.   usage: job_scale.sh N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
if [ "$N" != "1" ]; then
  timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.log 2>&1; echo "bench$N rc=$?"
  grep metric gpurun_out/bench$N.log | cut -c1-300
  timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check$N.log 2>&1; echo "dist_check rc=$?"
  tail -2 gpurun_out/dist_check$N.log
fi
timeout 1200 $TR tests/scripts/c5_dist.py --size 32768 --dtypes f32,f64 > gpurun_out/c5_dist${N}_32768.log 2>&1; echo "c5 rc=$?"
grep config gpurun_out/c5_dist${N}_32768.log | cut -c1-900
