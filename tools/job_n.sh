#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/scripts/configs_run.py --oracle > gpurun_out/configs_run.log 2>&1; echo "rc=$?"
cat gpurun_out/configs_run.log | cut -c1-700
