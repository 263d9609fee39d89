#!/bin/bash
# round 2 (session 5): pipelined copies of the one-shot call (upload head start + download drain): parity tests, bench A/B
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 600 -x) > gpurun_out/r05_pytest_pipeline.log 2>&1; echo "pipeline rc=$?"; tail -12 gpurun_out/r05_pytest_pipeline.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r05_bench_1gpu_pipelined.json 2> gpurun_out/r05_bench_1gpu_pipelined.err; echo "bench rc=$?"; cat gpurun_out/r05_bench_1gpu_pipelined.json; tail -3 gpurun_out/r05_bench_1gpu_pipelined.err
