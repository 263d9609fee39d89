#!/bin/bash
# round 2 (session 5), single GPU: full regression on the restored tree, bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r05_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r05_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r05_bench_1gpu.json 2> gpurun_out/r05_bench_1gpu.err; echo "bench rc=$?"; cat gpurun_out/r05_bench_1gpu.json
