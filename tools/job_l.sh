#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; head -13 gpurun_out/e2e_breakdown.log
timeout 120 python tests/scripts/small_latency.py > gpurun_out/small_latency.log 2>&1; cat gpurun_out/small_latency.log
