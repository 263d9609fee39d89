"""Compare stream kernel vs oracle and print the error pattern."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from heightmap_interpolation_b200 import _capi as C, synthetic
from heightmap_interpolation_b200.solver import Solver, make_params
from oracle.fdpde_oracle import OracleInpainter

def run(method, rows, cols, tb, K, chunk, dtype=np.float64):
    img = synthetic.bathymetry((rows, cols), dtype=dtype); mask = synthetic.mask_random((rows, cols), 0.4, 3)
    img = synthetic.zero_unknown(img, mask)
    code, step = {"harmonic": (C.HMI_HARMONIC, 0.2), "ccst": (C.HMI_CCST, 0.01)}[method]
    outs = []
    for rep in range(3):
        p = make_params(rows, cols, dtype, code, orientation=1, dt=step, tension=0.3, kernel=C.HMI_KERNEL_STREAM, temporal_block=tb)
        with Solver(p) as s:
            if chunk: s.set_option("chunk_rows", chunk)
            s.set_problem_host(img, mask); s.sweeps(K); outs.append(s.get_result_host())
    ref = OracleInpainter(method, convolver="opencv", update_step_size=step, tension=0.3, max_iters=K, term_thres=1e-300, term_check_iters=10**9).inpaint_grid(img.copy(), mask)
    d = np.abs(outs[0] - ref)
    bad = d > 1e-9 * (ref.max() - ref.min())
    print(method, rows, cols, "tb", tb, "K", K, "chunk", chunk, "maxerr", d.max(), "nbad", bad.sum(),
          "deterministic", np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2]))
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("  rows", np.unique(ys)[:20], "... cols", np.unique(xs)[:40])
        print("  rows mod 3 hist", np.bincount(ys % 3, minlength=3), " col hist by window pos", np.bincount((xs % 120) // 4, minlength=30)[:30])

run("harmonic", 96, 300, 2, 4, 0)
run("harmonic", 96, 300, 2, 5, 0)
run("harmonic", 211, 333, 2, 5, 37)
run("harmonic", 211, 333, 2, 4, 37)
run("harmonic", 211, 333, 2, 2, 37)
run("harmonic", 48, 64, 4, 8, 0)
run("harmonic", 48, 64, 4, 4, 0)
run("ccst", 211, 333, 2, 4, 37)
