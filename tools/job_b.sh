#!/bin/bash
# scratch GPU job: parity tests (per-test timeout), chunk policy A/B with geometry print, ncu captures, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
HMI_DEBUG=1 timeout 600 python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,4,6 --chunks=-1,0 --vecs 2 > gpurun_out/tune_b.log 2> gpurun_out/tune_b.err
sort -u gpurun_out/tune_b.err | grep "norm=0" > gpurun_out/tune_b_geom.log
timeout 300 python tools/tune.py --n 8192 --methods ccst,harmonic --dtypes f64 --kernels stream --tbs 3,6 --chunks=-1,0 --vecs 2 >> gpurun_out/tune_b.log 2>&1
timeout 300 python tools/tune.py --n 2048 --methods ccst,harmonic --dtypes f64 --kernels stream --tbs 3,6 --chunks=-1,0 --vecs 2 >> gpurun_out/tune_b.log 2>&1
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log
# ncu: full captures of the hot kernels (after the un-profiled runs above exited)
for cfg in "ccst f64 0" "tv f64 0" "amle f64 0" "tv f32 0"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream -s 2 -c 2 -f -o gpurun_out/prof_$1_$2 \
     python tools/prof_one.py --method $1 --dtype $2 --n 4096 --sweeps 12 > gpurun_out/ncu_$1_$2.log 2>&1
done
