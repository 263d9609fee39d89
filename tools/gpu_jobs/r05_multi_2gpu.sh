#!/bin/bash
# round 2 (session 5), 2 GPUs: multi-GPU tests (device groups on real peers, one process per GPU over IPC links / NCCL), bench at N = 2
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900) > gpurun_out/r05_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r05_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r05_bench_2gpu.json 2> gpurun_out/r05_bench_2gpu.err; echo "bench rc=$?"; grep '^{' gpurun_out/r05_bench_2gpu.json | cut -c1-1500; tail -3 gpurun_out/r05_bench_2gpu.err
