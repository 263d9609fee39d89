"""Sweep kernel variants / temporal-block depth / chunk height and print GLUPS (device-resident).

    python tools/tune.py [--n 4096] [--methods ccst,harmonic] [--dtypes f64,f32]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heightmap_interpolation_b200 import _capi as C                      # noqa: E402
from heightmap_interpolation_b200 import synthetic                       # noqa: E402
from heightmap_interpolation_b200.solver import Solver, make_params      # noqa: E402

M = {"harmonic": (C.HMI_HARMONIC, 0.2), "ccst": (C.HMI_CCST, 0.01), "tv": (C.HMI_TV, 0.225),
     "amle": (C.HMI_AMLE, 0.01)}


def time_sweeps(s, n, reps=3):
    s.sweeps(n)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.sweeps(n)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--methods", default="ccst,harmonic")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--kernels", default="stream,generic")
    ap.add_argument("--tbs", default="1,2,3,4,6,8")
    ap.add_argument("--chunks", default="0,32,64,128,256")
    ap.add_argument("--sweeps", type=int, default=48)
    ap.add_argument("--mask", default="tracks")
    ap.add_argument("--vecs", default="2,4")
    ap.add_argument("--tma", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0, help="rows of the grid if not square (a row slab: --rows 4096 --n 32768)")
    ap.add_argument("--border", default="reflect101", choices=["reflect101", "symmetric", "zero"])
    ap.add_argument("--orientation", type=int, default=1)
    a = ap.parse_args()
    n = a.n
    rows = a.rows or n
    for dt in a.dtypes.split(","):
        dtype = np.float64 if dt == "f64" else np.float32
        img = synthetic.bathymetry((rows, n), dtype=dtype)
        mask = synthetic.mask_tracks((rows, n)) if a.mask == "tracks" else synthetic.mask_random((rows, n), 0.4, 3)
        img = synthetic.zero_unknown(img, mask)
        bpl = 17 if dt == "f64" else 9
        for method in a.methods.split(","):
            code, step = M[method]
            for kernel in a.kernels.split(","):
                tbs = [int(t) for t in a.tbs.split(",")] if kernel == "stream" else [1]
                if method in ("tv", "amle"):
                    tbs = [t for t in tbs if t <= 3] or [1]
                for tb in tbs:
                    if kernel == "stream" and method == "ccst" and tb > 4:
                        continue
                    chunks = [int(c) for c in a.chunks.split(",")] if kernel == "stream" else [0]
                    for ch, sk in [(c, k) for c in chunks for k in ([int(x) for x in a.vecs.split(",")] if (kernel == "stream" and dt == "f64") else [0])]:
                        p = make_params(rows, n, dtype, code, orientation=a.orientation, dt=step, tension=0.3, epsilon=1.0,
                                        border={"reflect101": C.HMI_BORDER_REFLECT101, "symmetric": C.HMI_BORDER_SYMMETRIC,
                                                "zero": C.HMI_BORDER_ZERO}[a.border],
                                        kernel=C.KERNELS[kernel], temporal_block=tb)
                        with Solver(p) as s:
                            if ch:
                                s.set_option("chunk_rows", ch)
                            if kernel == "stream" and method not in ("tv", "amle"):
                                s.set_option("vec", sk)
                                s.set_option("tma", a.tma)
                            s.set_problem_host(img, mask)
                            nsw = a.sweeps if kernel == "stream" else max(4, a.sweeps // 4)
                            nsw = (nsw // tb) * tb
                            ms = time_sweeps(s, nsw)
                        g = rows * n * nsw / (ms * 1e-3) / 1e9
                        print(("" if a.border == "reflect101" else a.border + " border ") + "%-8s %s %-7s tb=%d chunk=%-4d vec=%d tma=%d  %8.1f GLUPS  %6.1f%% of HBM roofline (%.0f GB/s algorithmic)"
                              % (method, dt, kernel, tb, ch, sk, a.tma, g, 100 * g * bpl / 6550.4, g * bpl), flush=True)


if __name__ == "__main__":
    main()
