"""Run a few sweeps of one configuration (for ncu captures).
    python tools/prof_one.py --method ccst --dtype f64 --tb 2 --n 4096 --sweeps 12 [--chunk 128]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heightmap_interpolation_b200 import _capi as C                      # noqa: E402
from heightmap_interpolation_b200 import synthetic                       # noqa: E402
from heightmap_interpolation_b200.solver import Solver, make_params      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--method", default="ccst")
ap.add_argument("--dtype", default="f64")
ap.add_argument("--kernel", default="stream")
ap.add_argument("--tb", type=int, default=0)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--sweeps", type=int, default=12)
ap.add_argument("--mask", default="tracks")
ap.add_argument("--vec", type=int, default=-1)
ap.add_argument("--tma", type=int, default=0)
a = ap.parse_args()
M = {"harmonic": (C.HMI_HARMONIC, 0.2), "ccst": (C.HMI_CCST, 0.01), "tv": (C.HMI_TV, 0.225), "amle": (C.HMI_AMLE, 0.01)}
dtype = np.float64 if a.dtype == "f64" else np.float32
img = synthetic.bathymetry(a.n, dtype=dtype)
mask = synthetic.mask_tracks(a.n) if a.mask == "tracks" else synthetic.mask_random(a.n, 0.4, 3)
img = synthetic.zero_unknown(img, mask)
code, step = M[a.method]
p = make_params(a.n, a.n, dtype, code, orientation=1, dt=step, tension=0.3, epsilon=1.0,
                kernel=C.KERNELS[a.kernel], temporal_block=a.tb)
with Solver(p) as s:
    if a.chunk:
        s.set_option("chunk_rows", a.chunk)
    if a.vec >= 0:
        s.set_option("vec", a.vec)
    if a.tma:
        s.set_option("tma", a.tma)
    s.set_problem_host(img, mask)
    nm = s.sweeps_norm(a.sweeps)
    print(s.kernel_name, s.temporal_block, s.launch_count, nm.sum_d2, nm.sum_f2)
