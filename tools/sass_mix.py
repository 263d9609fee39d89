"""Instruction mix of the loops of one kernel in an object file / library (no GPU needed).

    python tools/sass_mix.py file.o 'stream_kernel<double, (int)2, (int)3, (int)1,' [--dump lo hi]
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
funcs = [(i, l.split(":")[1].strip()) for i, l in enumerate(txt) if "Function :" in l]
dem = subprocess.run(["cu++filt"] + [f[1] for f in funcs], capture_output=True, text=True).stdout.split("\n")
for k, ((i, n), d) in enumerate(zip(funcs, dem)):
    if pat not in d:
        continue
    end = funcs[k + 1][0] if k + 1 < len(funcs) else len(txt)
    ins = []
    for l in txt[i:end]:
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("==", d, len(ins), "instructions")
    if "--dump" in sys.argv:
        lo, hi = (int(x, 16) for x in sys.argv[sys.argv.index("--dump") + 1:][:2])
        for a, t in ins:
            if lo <= a <= hi:
                print("%05x %s" % (a, t))
        continue
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+.*?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            c = collections.Counter()
            for _, x in body:
                x = re.sub(r"^@!?U?P\w+\s+", "", x)
                c[x.split()[0].split(".")[0]] += 1
            f64 = c["DADD"] + c["DFMA"] + c["DMUL"]
            f32 = c["FADD"] + c["FFMA"] + c["FMUL"]
            print("loop %05x..%05x n=%d fp64=%d fp32=%d  %s" % (tgt, a, len(body), f64, f32, dict(c.most_common(14))))
