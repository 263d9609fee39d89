#!/bin/bash
# round 2 (session 5): the chunk picker's load-balance term (half a round of drain time per launch) against the old cost
# model, over methods, precisions and grid sizes (one-wave grids, slabs, many-wave grids)
mkdir -p gpurun_out
for b in 0 0.5; do
export HMI_PICK_ROUND_BIAS=$b
echo "== round bias $b"
for n in 2048 4096 8192; do
python tools/tune.py --n $n --methods harmonic --dtypes f64 --kernels stream --tbs 6 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods harmonic --dtypes f32 --kernels stream --tbs 8 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods ccst --dtypes f32 --kernels stream --tbs 4 --chunks 0 --vecs 2 --sweeps 96
python tools/tune.py --n $n --methods tv,amle --dtypes f64,f32 --kernels stream --tbs 2 --chunks 0 --vecs 2 --sweeps 96
done
python tools/tune.py --n 16384 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 48 --mask random
python tools/tune.py --n 16384 --methods tv --dtypes f32 --kernels stream --tbs 2 --chunks 0 --vecs 2 --sweeps 48 --mask random
done > gpurun_out/r05_tune_picker_round_bias.log 2>&1
python - <<'PY'
import re, collections
cur, rows = None, collections.OrderedDict()
n = None
for l in open("gpurun_out/r05_tune_picker_round_bias.log"):
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
