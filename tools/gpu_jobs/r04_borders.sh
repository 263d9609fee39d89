#!/bin/bash
# round 2 (session 4): zero-padded / symmetric borders on the streaming kernels: parity tests, then throughput vs the generic kernel
mkdir -p gpurun_out; nvidia-smi > /dev/null 2>&1; sleep 2
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "zero or 1d or border or generic or staging or stream_kernel_matches" ) > gpurun_out/r04_pytest_borders.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r04_pytest_borders.log | cut -c1-250
