#!/bin/bash
# ncu --set full of every hot kernel at 4096^2; only the text summaries come back (the reports stay in /tmp)
mkdir -p gpurun_out
for cfg in "harmonic f64" "ccst f32" "harmonic f32" "tv f64" "tv f32" "amle f64" "amle f32"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream -s 3 -c 1 -f -o /tmp/prof3_$1_$2 \
     python tools/prof_one.py --method $1 --dtype $2 --n 4096 --sweeps 24 > /tmp/ncu3_$1_$2.log 2>&1; echo "$cfg rc=$?"
  python tools/ncu_summary.py /tmp/prof3_$1_$2.ncu-rep > gpurun_out/ncu3_$1_$2_summary.txt 2>&1
  python tools/ncu_opmix.py /tmp/prof3_$1_$2.ncu-rep 1 2>&1 | head -16 >> gpurun_out/ncu3_$1_$2_summary.txt
done
