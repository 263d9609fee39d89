#!/bin/bash
# round 2 (session 5): ncu launch list of the bench command on the 32768^2 workload (kernel share of the step), and a
# full capture of the dominant kernel on the same workload, summarised on the box (the reports exceed gpurun_out's limit)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --sweeps 30 --no-cpu-baseline --no-extra > gpurun_out/r05_bench_short.json 2>/dev/null; echo "bench short rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r05_ncu_launches_bench_32768.csv python bench.py --steps 2 --warmup 3 --sweeps 30 --no-cpu-baseline --no-extra > gpurun_out/r05_ncu_l.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r05_ncu_launches_bench_32768.csv")) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.defaultdict(float); n = collections.Counter()
for r in rows:
    name = r[k].split("(")[0][:90]
    t[name] += float(r[v].replace(",", "")); n[name] += 1
tot = sum(t.values())
with open("gpurun_out/r05_ncu_launches_bench_32768_summary.txt", "w") as fh:
    for name in sorted(t, key=t.get, reverse=True):
        line = "%6.2f %%  %5d launches  %10.1f us each  %s" % (100 * t[name] / tot, n[name], t[name] / n[name] / 1e3, name)
        print(line); fh.write(line + "\n")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o /tmp/r05_full python tools/prof_one.py --n 32768 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random > /tmp/r05_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py /tmp/r05_full.ncu-rep > gpurun_out/r05_ncu_ccst_f64_k3_32768_summary.txt 2>&1
python tools/ncu_opmix.py /tmp/r05_full.ncu-rep 1 2>&1 | head -45 >> gpurun_out/r05_ncu_ccst_f64_k3_32768_summary.txt
head -24 gpurun_out/r05_ncu_ccst_f64_k3_32768_summary.txt
