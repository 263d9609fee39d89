"""Per-opcode executed-instruction mix and top stall sites from an ncu report's source page.
    python tools/ncu_opmix.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.reader(('"Kernel Name"' + blocks[which]).splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
si, ie, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, smp = Counter(), Counter()
lines = []
for r in rows[2:]:
    if len(r) <= ie:
        continue
    toks = r[si].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    n = int(r[ie])
    c[op.split(".")[0]] += n
    smp[op.split(".")[0]] += int(r[ss])
    lines.append((int(r[ss]), n, r[si].strip()))
tot, ts = sum(c.values()), sum(smp.values())
print("total warp instructions", tot, "samples", ts)
for k, v in c.most_common(22):
    print("%-10s %10d %5.1f%%   stall-samples %5.1f%%" % (k, v, 100 * v / tot, 100 * smp[k] / max(ts, 1)))
print("--- top stall sites")
for s, n, src in sorted(lines, reverse=True)[:18]:
    print("%6d %9d  %s" % (s, n, src[:90]))
