"""Multigrid end to end: harmonic / CCST inpainting of an n x n grid with large holes, mgs_levels levels.
    python tools/mg_bench.py [--n 4096] [--levels 4] [--dtype f32]
Prints the wall time of inpaint() (host arrays in, host array out) and the per-level iteration counts."""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heightmap_interpolation_b200 as hmi                        # noqa: E402
from heightmap_interpolation_b200 import synthetic                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--levels", type=int, default=4)
ap.add_argument("--dtype", default="f32")
a = ap.parse_args()
dtype = np.float32 if a.dtype == "f32" else np.float64
img0 = synthetic.bathymetry(a.n, dtype=dtype)
mask = synthetic.mask_holes(a.n, 64, max(8, a.n // 64), seed=1)
img0 = synthetic.zero_unknown(img0, mask)
for name, cls, opts in (("harmonic", hmi.SobolevInpainter, dict(update_step_size=0.2)),
                        ("ccst", hmi.CCSTInpainter, dict(update_step_size=0.01, tension=0.3))):
    for levels in (1, a.levels):
        inp = cls(term_criteria="absolute_percent", term_thres=1e-5, term_check_iters=1000, max_iters=200000,
                  init_with="nearest", convolver="opencv", mgs_levels=levels, mgs_min_res=100, **opts)
        best = 1e9
        for rep in range(2):
            img = img0.copy()
            t = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                out = inp.inpaint(img, mask)
            best = min(best, time.perf_counter() - t)
            inp.term_thres = 1e-5          # absolute_percent rescales it in place (as the reference does)
        print("%-8s %s %dx%d mgs_levels=%d  inpaint() %.3f s   levels (rows, cols, iterations): %s  converged=%s"
              % (name, a.dtype, a.n, a.n, levels, best, inp.history_, inp.converged_), flush=True)
