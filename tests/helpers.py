"""Shared helpers for the parity tests."""
import ast

import numpy as np

# Per-method options used by tests/golden/make_golden.py (CLI defaults of the reference)
METHOD_OPTS = {
    "harmonic": {"update_step_size": 0.2},
    "ccst": {"update_step_size": 0.01, "tension": 0.3},
    "tv": {"update_step_size": 0.225, "epsilon": 1.0},
    "amle": {"update_step_size": 0.01},
}

# Parity tolerances (BASELINE.json north_star): filled cells within 1e-10 (fp64) / 1e-5 (fp32),
# measured as max|a-b| / (max(ref) - min(ref)) (SURVEY.md section 8d "Parity protocol").
TOL = {"f64": 1e-10, "f32": 1e-5}
DTYPES = {"f64": np.float64, "f32": np.float32}


def range_err(a, ref):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    rng = float(ref.max() - ref.min())
    if rng == 0.0:
        rng = 1.0
    return float(np.abs(a - ref).max()) / rng


def parse_converged_meta(golden):
    """Rows of the converged.npz 'meta' table -> list of dicts."""
    out = []
    for row in golden["meta"]:
        key, method, conv, dt, crit, thres, T, extra, upd, diff, term, final_thres = str(row).split("|")
        out.append(dict(key=key, method=method, convolver=conv, dtype=dt, criteria=crit,
                        thres=float(thres), T=int(T), extra=dict(ast.literal_eval(extra)),
                        updates=int(upd), diff=float(diff), converged=bool(int(term)),
                        final_thres=float(final_thres)))
    return out
