#!/bin/bash
# round 2 (session 5): regression after the picker's load-balance term, old vs new cost model, bench line
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900) > gpurun_out/r05_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r05_pytest_gpu_final.log
for b in 0 0.5; do
export HMI_PICK_ROUND_BIAS=$b
echo "== round bias $b"
for n in 2048 4096 8192; do
python tools/tune.py --n $n --methods harmonic --dtypes f32 --kernels stream --tbs 8 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods tv,amle --dtypes f64,f32 --kernels stream --tbs 2 --chunks 0 --vecs 2 --sweeps 96
done
done > gpurun_out/r05_tune_picker_round_bias_min48.log 2>&1
python - <<'PY'
import re, collections
cur, rows = None, collections.OrderedDict()
for l in open("gpurun_out/r05_tune_picker_round_bias_min48.log"):
    if l.startswith("== round bias"):
        cur = l.split()[-1]; idx = 0; continue
    m = re.match(r"(\w+)\s+(f\d+) stream\s+tb=(\d).*?([\d.]+) GLUPS", l)
    if m:
        rows.setdefault(idx, {})[cur] = (m.group(1), m.group(2), m.group(3), float(m.group(4))); idx += 1
for i, d in rows.items():
    a, b = d.get("0"), d.get("0.5")
    if a and b:
        print("%-8s %s tb=%s  bias 0: %7.1f   bias 0.5: %7.1f   %+5.1f %%" % (a[0], a[1], a[2], a[3], b[3], 100 * (b[3] / a[3] - 1)))
PY
unset HMI_PICK_ROUND_BIAS
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r05_bench_1gpu_final.json 2> gpurun_out/r05_bench_1gpu_final.err; echo "bench rc=$?"; cat gpurun_out/r05_bench_1gpu_final.json
