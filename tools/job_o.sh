#!/bin/bash
# 2-GPU experiment: halo exchange overhead on small fp32 slabs, NCCL SM kernels vs copy-engine P2P
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
T1="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29562"
timeout 300 $T1 tests/scripts/c5_dist.py --size 8192 --dtypes f32,f64 > gpurun_out/exch_1gpu.log 2>&1
timeout 300 $TR tests/scripts/c5_dist.py --size 8192 --dtypes f32,f64 > gpurun_out/exch_2gpu_nccl.log 2>&1
NCCL_P2P_USE_CUDA_MEMCPY=1 timeout 300 $TR tests/scripts/c5_dist.py --size 8192 --dtypes f32,f64 > gpurun_out/exch_2gpu_ce.log 2>&1
for f in exch_1gpu exch_2gpu_nccl exch_2gpu_ce; do echo "== $f"; grep config gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], 'iters', d['iterations'], 'resident %.0f GLUPS'%d['glups_resident'], 'solve %.3f s'%d['solve_s'])
"; done
