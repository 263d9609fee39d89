#!/bin/bash
# round 2: halo links / device groups / IPC slabs on 2 GPUs, then the bench at N = 2 (and N = 1 for the block hashes)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q --timeout 600 > gpurun_out/r03_test_multi.log 2>&1; echo "multi tests rc=$?"; tail -15 gpurun_out/r03_test_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r03_bench_2gpu.json 2> gpurun_out/r03_bench_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r03_bench_2gpu.err; cat gpurun_out/r03_bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --halo nccl > gpurun_out/r03_bench_2gpu_nccl.json 2> gpurun_out/r03_bench_2gpu_nccl.err; echo "bench2 nccl rc=$?"; tail -5 gpurun_out/r03_bench_2gpu_nccl.err; cat gpurun_out/r03_bench_2gpu_nccl.json
