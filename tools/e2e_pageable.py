"""End-to-end time of the public call on PAGEABLE NumPy arrays (what a user of the reference passes), BASELINE
config 5 solved to convergence, with and without the copies overlapped with the solve.
    python tools/e2e_pageable.py [--n 32768]
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heightmap_interpolation_b200 as hmi                     # noqa: E402
from heightmap_interpolation_b200 import _capi as C            # noqa: E402
from heightmap_interpolation_b200 import synthetic             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
a = ap.parse_args()
n = a.n
img = np.empty((n, n))
mask = np.empty((n, n), dtype=bool)
for r in range(0, n, 2048):
    mask[r:r + 2048] = synthetic.mask_random((n, n), 0.40, seed=3, row_range=(r, r + 2048))
    img[r:r + 2048] = synthetic.bathymetry((n, n), seed=7, row_range=(r, r + 2048))
img[~mask] = 0
out = np.empty_like(img)
C.lib().hmi_host_prefault(out.ctypes.data, out.nbytes)
cases = [("converged (relative 1e-5, check every 50)", dict(term_thres=1e-5, term_check_iters=50, max_iters=100000)),
         ("300 sweeps, no check after the first", dict(term_thres=1e-300, term_check_iters=1000, max_iters=300))]
for name, kw in cases:
    for mode in ("0", "1", "0", "1"):
        os.environ["HMI_UPLOAD_PIPELINE"] = mode
        inp = hmi.CCSTInpainter(update_step_size=0.01, tension=0.3, convolver="opencv", init_with="zeros", **kw)
        work = img
        t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            inp.inpaint(work, mask, out=out)
        dt = time.perf_counter() - t
        print("%-44s copies %s: %6.3f s, %4d sweeps, %6.1f GLUPS end to end" %
              (name, "overlapped" if mode == "1" else "plain     ", dt, inp.iterations_, n * n * inp.iterations_ / dt / 1e9), flush=True)
