#!/bin/bash
# final single-GPU measurements: smoke, bench (both arms), ncu launch list of the bench command, full capture of the top kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --sweeps 100 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 2 -f -o gpurun_out/prof_bench python bench.py --steps 1 --warmup 3 --sweeps 60 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc=$?"
