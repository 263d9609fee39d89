#!/bin/bash
# round 2 (session 5): TV / AMLE after the per-kernel minimum chunk height (fp32: 96 rows): parity tests and old vs new cost model
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "tv or amle or nl or mid_sized") 2>&1 | tail -2
for b in 0 0.5; do
export HMI_PICK_ROUND_BIAS=$b
echo "== round bias $b"
for n in 4096 8192; do
python tools/tune.py --n $n --methods tv,amle --dtypes f64,f32 --kernels stream --tbs 2 --chunks 0 --vecs 2 --sweeps 96
done
python tools/tune.py --n 16384 --methods tv,amle --dtypes f32 --kernels stream --tbs 2 --chunks 0 --vecs 2 --sweeps 48 --mask random
done > gpurun_out/r05_tune_picker_round_bias_nl_min96.log 2>&1
python - <<'PY'
import re, collections
cur, rows = None, collections.OrderedDict()
for l in open("gpurun_out/r05_tune_picker_round_bias_nl_min96.log"):
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
