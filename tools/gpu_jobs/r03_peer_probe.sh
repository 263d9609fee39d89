#!/bin/bash
# round 2: can halo rows be pushed by the copy engines with flags in peer memory (in-process and CUDA IPC)?
mkdir -p gpurun_out
{
nvidia-smi topo -m
for mode in inproc ipc; do for fl in memops kernel; do
  timeout 60 tools/probes/build/peer_probe $mode $fl; echo "rc=$?"
done; done
} > gpurun_out/r03_peer_probe.log 2>&1
cat gpurun_out/r03_peer_probe.log
