#!/bin/bash
# round 2 (session 4), single GPU: whole GPU suite at HEAD, smoke, bench (both arms), ncu launch list of the bench, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r04_gpu.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r04_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r04_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r04_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r04_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/r04_bench_ref.json 2> gpurun_out/r04_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r04_bench_ref.json
( time timeout 900 python bench.py ) > gpurun_out/r04_bench_1gpu.json 2> gpurun_out/r04_bench_1gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r04_bench_1gpu.err; cat gpurun_out/r04_bench_1gpu.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r04_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --sweeps 60 --no-cpu-baseline --no-extra > gpurun_out/r04_ncu_l.log 2>&1; echo "ncu list rc=$?"
for tma in 0 2; do
  for n in 4096 16384; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/r04_ccst_f64_k3_${n}_tma${tma} python tools/prof_one.py --n $n --method ccst --dtype f64 --tb 3 --sweeps 12 --tma $tma > gpurun_out/r04_ncu_${n}_${tma}.log 2>&1; echo "ncu $n tma=$tma rc=$?"
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o gpurun_out/r04_ccst_f64_k3_32768 python tools/prof_one.py --n 32768 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random > gpurun_out/r04_ncu_32768.log 2>&1; echo "ncu 32768 rc=$?"
ls -la gpurun_out/*.ncu-rep
