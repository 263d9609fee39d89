#!/bin/bash
# scratch N-GPU job: slab parity (TV with two fused sweeps), BASELINE config 4 (16384^2 fp64 biharmonic + harmonic), bench at N
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check$N.log 2>&1; echo "dist_check rc=$?"
tail -6 gpurun_out/dist_check$N.log
timeout 900 $TR tests/scripts/c5_dist.py --size 16384 --dtypes f64 --methods harmonic,ccst --tension 0 --seed 2 > gpurun_out/c4_dist${N}_16384.log 2>&1; echo "c4 rc=$?"
grep config gpurun_out/c4_dist${N}_16384.log | cut -c1-800
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.log 2>&1; echo "bench$N rc=$?"
grep metric gpurun_out/bench$N.log | cut -c1-300
