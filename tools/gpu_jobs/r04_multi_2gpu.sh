#!/bin/bash
# round 2 (session 4), 2 GPUs: multi-GPU tests (device groups, halo links, one process per GPU over IPC + NCCL), bench at N = 2
mkdir -p gpurun_out; nvidia-smi > /dev/null 2>&1; sleep 2
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 ) > gpurun_out/r04_pytest_multi_2gpu.log 2>&1; echo "multi tests rc=$?"; tail -8 gpurun_out/r04_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r04_bench_2gpu.json 2> gpurun_out/r04_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r04_bench_2gpu.err; cat gpurun_out/r04_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r04_bench_2gpu_ref.json 2> gpurun_out/r04_bench_2gpu_ref.err; echo "ref2 rc=$?"; cat gpurun_out/r04_bench_2gpu_ref.json | cut -c1-300
