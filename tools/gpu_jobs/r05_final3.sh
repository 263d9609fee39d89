#!/bin/bash
# round 2 (session 5): last regression of the tree as committed: every GPU test, smoke, bench line
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900) > gpurun_out/r05_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r05_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r05_bench_1gpu_final.json 2> gpurun_out/r05_bench_1gpu_final.err; echo "bench rc=$?"; cat gpurun_out/r05_bench_1gpu_final.json
