#!/bin/bash
# round 2 (session 5), N GPUs: bench (strong scaling of the 32768^2 grid, peer halo links) and BASELINE config 5 solved to convergence
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r05_bench_${N}gpu.json 2> gpurun_out/r05_bench_${N}gpu.err; echo "bench$N rc=$?"
grep '^{' gpurun_out/r05_bench_${N}gpu.json | cut -c1-1200; tail -3 gpurun_out/r05_bench_${N}gpu.err | cut -c1-300
timeout 900 $TR tests/scripts/c5_dist.py --size 32768 --dtypes f32,f64 > gpurun_out/r05_c5_32768_${N}gpu.log 2>&1; echo "c5 rc=$?"
grep config gpurun_out/r05_c5_32768_${N}gpu.log | cut -c1-700
