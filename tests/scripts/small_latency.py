"""Latency of inpaint() on small grids (the reference CLI inpaints one work area at a time, often small):
BASELINE config 1 (harmonic, 512^2 Gaussian hill, 30 % random mask, fp64, relative 1e-5, check every 10)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import synthetic
from oracle.fdpde_oracle import OracleInpainter

for n in (128, 512, 1024, 2048):
    img = synthetic.gaussian_hill(n); mask = synthetic.mask_random(n, 0.30, seed=0); img = synthetic.zero_unknown(img, mask)
    opts = dict(update_step_size=0.2, term_thres=1e-5, term_check_iters=10, convolver="opencv")
    inp = hmi.SobolevInpainter(**opts)
    inp.inpaint(img.copy(), mask)
    ts = []
    for _ in range(20):
        a = img.copy()
        t = time.perf_counter(); out = inp.inpaint(a, mask); ts.append(time.perf_counter() - t)
    o = OracleInpainter("harmonic", **opts)
    t = time.perf_counter(); ref = o.inpaint(img.copy(), mask); tc = time.perf_counter() - t
    err = float(np.abs(out - ref).max()) / float(ref.max() - ref.min())
    print("harmonic %4d^2 fp64: %d iterations (oracle %d), inpaint() median %.3f ms (min %.3f), CPU oracle %.1f ms, err %.1e"
          % (n, inp.iterations_, o.updates_done, 1e3 * float(np.median(ts)), 1e3 * min(ts), 1e3 * tc, err), flush=True)
