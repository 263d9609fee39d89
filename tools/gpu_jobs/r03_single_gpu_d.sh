#!/bin/bash
# round 2, single GPU: new tests (device pyramid, CLI areas), ncu captures (bench kernel at 32768^2; cp.async vs tensor-map staging)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hooks.py -m gpu -q --timeout 900 > gpurun_out/r03_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r03_pytest_gpu_d.log
timeout 600 python tools/prof_one.py --n 4096 --method ccst --dtype f64 --tb 3 --sweeps 12 > gpurun_out/prof_plain.log 2>&1; echo "plain rc=$?"
for tma in 0 2; do
  for n in 4096 16384; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/r03_ccst_f64_k3_${n}_tma${tma} python tools/prof_one.py --n $n --method ccst --dtype f64 --tb 3 --sweeps 12 --tma $tma > gpurun_out/ncu_${n}_${tma}.log 2>&1; echo "ncu $n tma=$tma rc=$?"
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/r03_ccst_f64_k3_32768 python tools/prof_one.py --n 32768 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random > gpurun_out/ncu_32768.log 2>&1; echo "ncu 32768 rc=$?"
ls -la gpurun_out/*.ncu-rep
