#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for n in 4096 8192 16384; do
HMI_DEBUG=1 timeout 600 python tools/tune.py --n $n --methods tv,amle --dtypes f64,f32 --kernels stream --chunks=-1,0 >> gpurun_out/tune_nl.log 2>> gpurun_out/tune_nl.err
done
sort -u gpurun_out/tune_nl.err | grep "norm=0" > gpurun_out/tune_nl_geom.log
timeout 600 python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f32 --kernels stream --tbs 3,4,6,8 --chunks=0 --vecs 2 > gpurun_out/tune_f32_16k.log 2>&1
for cfg in "tv f64" "amle f64" "tv f32"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream -s 2 -c 1 -f -o gpurun_out/prof2_$1_$2 \
     python tools/prof_one.py --method $1 --dtype $2 --n 4096 --sweeps 12 > gpurun_out/ncu2_$1_$2.log 2>&1
done
