"""Cost of one temporal block on a row band (the launches of the pipelined copies, csrc/hmi_capi.cu) against its
share of a whole-grid launch, for several chunk heights.
    python tools/band_launch.py [--rows 512] [--cols 32768]
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

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=512)
ap.add_argument("--cols", type=int, default=32768)
ap.add_argument("--chunks", default="0,64,86,103,128,171,256,512")
a = ap.parse_args()
for rows in (a.rows, 16 * a.rows):
    mask = synthetic.mask_random((rows, a.cols), 0.4, seed=3)
    img = synthetic.bathymetry((rows, a.cols))
    img[~mask] = 0
    for ch in [int(c) for c in a.chunks.split(",")]:
        if ch > rows:
            continue
        p = make_params(rows, a.cols, np.float64, C.HMI_CCST, orientation=1, dt=0.01, tension=0.3, temporal_block=3)
        with Solver(p) as s:
            if ch:
                s.set_option("chunk_rows", ch)
            s.set_problem_host(img, mask)
            s.sweeps(30)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.sweeps(300)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 100
        print("rows %5d chunk %4d: %7.1f us per 3-sweep launch, %6.1f GLUPS" % (rows, ch, us, rows * a.cols * 3 / us / 1e3), flush=True)
