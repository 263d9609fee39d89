"""Where does the end-to-end inpaint() call spend its time?  (4096^2 fp64 CCST, 1001 sweeps)"""
import os, sys, time, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import _capi as C, synthetic
from heightmap_interpolation_b200.solver import Solver, make_params
from heightmap_interpolation_b200.inpainting.initializer import Initializer

n = 4096
img = synthetic.bathymetry(n); mask = synthetic.mask_tracks(n); img = synthetic.zero_unknown(img, mask)
pin_img = torch.from_numpy(img).pin_memory().numpy(); pin_mask = torch.from_numpy(mask).pin_memory().numpy()
pin_out = torch.empty((n, n), dtype=torch.float64).pin_memory().numpy()
def T(label, f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    print("%-40s %8.2f ms" % (label, best * 1e3)); return r
T("host zero fill (Initializer)", lambda: Initializer("zeros").initialize(pin_img, pin_mask))
p = make_params(n, n, np.float64, C.HMI_CCST, orientation=1, dt=0.01, tension=0.3, term_check_iters=1000, max_iters=1001, term_thres=1e-300)
s = T("hmi_create", lambda: Solver(p), reps=1)
T("set_problem_host pinned", lambda: s.set_problem_host(pin_img, pin_mask))
T("set_problem_host pageable", lambda: s.set_problem_host(img, mask))
T("run 1001 sweeps", lambda: s.run())
T("get_result_host pinned", lambda: s.get_result_host(pin_out))
pg = np.empty((n, n)); T("get_result_host pageable (touched)", lambda: s.get_result_host(pg))
T("get_result_host pageable (fresh)", lambda: s.get_result_host(None), reps=1)
T("destroy", lambda: s.close(), reps=1)
def cd():
    s2 = Solver(p); s2.close()
T("create+destroy (cached blocks)", cd, reps=5)
import cProfile, pstats
def full(out):
    inp = hmi.CCSTInpainter(update_step_size=0.01, tension=0.3, convolver="opencv", term_check_iters=1000, max_iters=1001, term_thres=1e-300)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        return inp.inpaint(pin_img, pin_mask, out=out)
T("inpaint() pinned in/out", lambda: full(pin_out))
T("inpaint() pinned in, fresh out", lambda: full(None))

pr = cProfile.Profile(); pr.enable(); full(pin_out); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
