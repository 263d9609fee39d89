"""BASELINE config 5 on one GPU: 32768 x 32768, 40 % missing, CCST and harmonic, fp32 and fp64.
Device-resident throughput + size-independent checks (known cells exact, corner crop vs oracle)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from heightmap_interpolation_b200 import _capi as C, synthetic
from heightmap_interpolation_b200.solver import Solver, make_params
from oracle.fdpde_oracle import OracleInpainter

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
for dt in (np.float32, np.float64):
    t0 = time.time()
    img = synthetic.bathymetry(n, seed=7, dtype=dt)
    mask = synthetic.mask_random(n, 0.40, seed=3)
    img = synthetic.zero_unknown(img, mask)
    print("generated %dx%d %s in %.1f s" % (n, n, np.dtype(dt).name, time.time() - t0), flush=True)
    for method, code, step in (("ccst", C.HMI_CCST, 0.01), ("harmonic", C.HMI_HARMONIC, 0.2)):
        p = make_params(n, n, dt, code, orientation=1, dt=step, tension=0.3)
        with Solver(p) as s:
            s.set_problem_host(img, mask)
            K = 2 * s.temporal_block
            s.sweeps(K); torch.cuda.synchronize()
            c = 64; m = 2 * K + 2 if method == "ccst" else K + 2
            corner = np.empty((c, n), dtype=dt)
            full_rows = s.get_result_host()[:c]
            ref = OracleInpainter(method, convolver="opencv", update_step_size=step, tension=0.3, max_iters=K,
                                  term_thres=1e-300, term_check_iters=10 ** 9).inpaint_grid(
                                      img[:c + m, :c + m].copy(), mask[:c + m, :c + m])
            err = float(np.abs(full_rows[:c, :c].astype(np.float64) - ref[:c, :c]).max()) / float(ref.max() - ref.min())
            known_ok = bool(np.array_equal(full_rows[mask[:c]], img[:c][mask[:c]]))
            reps = 6 * s.temporal_block
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.sweeps(reps); e1.record(); torch.cuda.synchronize()
            g = n * n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
            bpl = 17 if dt == np.float64 else 9
            print("%-8s %s %dx%d tb=%d  %8.1f GLUPS  %.0f%% of HBM roofline   corner err %.1e  known exact %s"
                  % (method, np.dtype(dt).name, n, n, s.temporal_block, g, 100 * g * bpl / 6550.4, err, known_ok), flush=True)
