"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports the reference package from /root/reference with an empty ``matplotlib`` stub (it is imported
at module top but only used under ``debug_dir``) and with ``sobolev_inpainter`` imported first
(circular import otherwise).  Writes small ``.npz`` fixtures next to this
script.  The reference never returns its iteration count, so ``term_check`` is wrapped to count
calls (updates = (n_calls - 1) * T + 1 on convergence; fd_pde_inpainter.py:322-351).
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import heightmap_interpolation.inpainting.sobolev_inpainter as sob
    from heightmap_interpolation.inpainting.ccst_inpainter import CCSTInpainter
    from heightmap_interpolation.inpainting.tv_inpainter import TVInpainter
    from heightmap_interpolation.inpainting.amle_inpainter import AMLEInpainter
    return {"harmonic": sob.SobolevInpainter, "ccst": CCSTInpainter, "tv": TVInpainter,
            "amle": AMLEInpainter}


# Per-method options used throughout (CLI defaults, apps/apps_common.py:126-147)
METHOD_OPTS = {
    "harmonic": {"update_step_size": 0.2},
    "ccst": {"update_step_size": 0.01, "tension": 0.3},
    "tv": {"update_step_size": 0.225, "epsilon": 1.0},
    "amle": {"update_step_size": 0.01},
}


def surface(rows, cols, seed, dtype):
    """Smooth hill + ripples + a little noise, unknown cells zeroed later."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:rows, 0:cols].astype(np.float64)
    s = min(rows, cols) / 5.0
    z = 100.0 * np.exp(-((x - cols / 2.3) ** 2 + (y - rows / 1.9) ** 2) / (2 * s * s))
    z += 7.0 * np.sin(x / 3.1) * np.cos(y / 4.3) + 0.5 * rng.standard_normal((rows, cols))
    return z.astype(dtype)


def edge_mask(rows, cols, p, seed):
    """Random mask with unknown cells forced on all four corners and along each edge."""
    rng = np.random.default_rng(seed)
    m = rng.random((rows, cols)) > p
    m[0, 0] = m[0, -1] = m[-1, 0] = m[-1, -1] = False
    m[0, 3:9] = False
    m[-1, 5:12] = False
    m[4:10, 0] = False
    m[7:15, -1] = False
    m[1, 1] = False
    m[10:14, 20:25] = False  # a small contiguous hole
    m[20, 30] = True
    return m


def run_reference(cls, image, mask, **opts):
    """Run one reference solve; returns (result, updates_done, last_diff, converged)."""
    inp = cls(**opts)
    calls = {"n": 0, "diff": None, "term": False}
    orig = inp.term_check

    def counting(f, fnew):
        t, d = orig(f, fnew)
        calls["n"] += 1
        calls["diff"] = float(d)
        calls["term"] = bool(t)
        return t, d

    inp.term_check = counting
    img = image.copy()
    out = inp.inpaint(img, mask)
    T = opts.get("term_check_iters", 1000)
    if calls["term"]:
        updates = (calls["n"] - 1) * T + 1
    else:
        updates = int(opts.get("max_iters", 1e8))
    return out, updates, calls["diff"], calls["term"], img, inp


def main():
    classes = load_reference()
    import heightmap_interpolation.inpainting.differential as diff  # noqa: F401 (import check)

    # ---------------------------------------------------------------- A. single step_fun outputs
    out = {}
    rng = np.random.default_rng(123)
    field64 = rng.standard_normal((37, 53)) * 10.0
    work = np.ones((37, 53), dtype=bool)
    out["field"] = field64
    for method, cls in classes.items():
        for conv in ("opencv", "masked", "scipy-ndimage", "scipy-signal"):
            for dt_name, dtype in (("f64", np.float64), ("f32", np.float32)):
                inp = cls(convolver=conv, **METHOD_OPTS[method])
                f = field64.astype(dtype)
                step = inp.step_fun(f.copy(), work if conv.startswith("masked") else None)
                out["step__%s__%s__%s" % (method, conv, dt_name)] = np.asarray(step)
    np.savez_compressed(os.path.join(HERE, "step_fun.npz"), **out)

    # ---------------------------------------------------------------- B. K-sweep states
    rows, cols = 48, 64
    mask = edge_mask(rows, cols, 0.35, 11)
    out = {"mask": mask}
    for dt_name, dtype in (("f64", np.float64), ("f32", np.float32)):
        img = surface(rows, cols, 5, dtype)
        img[~mask] = 0
        out["image__" + dt_name] = img
        for method, cls in classes.items():
            for conv in ("opencv", "masked"):
                for K in (1, 10, 60):
                    res, upd, _, _, _, _ = run_reference(
                        cls, img, mask, convolver=conv, max_iters=K, term_thres=1e-300,
                        term_check_iters=10 ** 9, init_with="zeros", **METHOD_OPTS[method])
                    assert upd == K and res.dtype == dtype
                    out["sweeps__%s__%s__%s__%d" % (method, conv, dt_name, K)] = res
    # other border rules (f64, 10 sweeps)
    img = out["image__f64"]
    for method in ("harmonic", "ccst", "tv", "amle"):
        for conv in ("scipy-ndimage", "scipy-signal"):
            res, _, _, _, _, _ = run_reference(
                classes[method], img, mask, convolver=conv, max_iters=10, term_thres=1e-300,
                term_check_iters=10 ** 9, **METHOD_OPTS[method])
            out["sweeps__%s__%s__f64__10" % (method, conv)] = res
    # CCST tension end points
    for t in (0.0, 1.0):
        res, _, _, _, _, _ = run_reference(
            classes["ccst"], img, mask, convolver="opencv", max_iters=10, term_thres=1e-300,
            term_check_iters=10 ** 9, update_step_size=0.01, tension=t)
        out["sweeps__ccst_t%d__opencv__f64__10" % int(t)] = res
    # over-relaxation (f64 only)
    res, _, _, _, _, _ = run_reference(
        classes["harmonic"], img, mask, convolver="opencv", max_iters=10, term_thres=1e-300,
        term_check_iters=10 ** 9, update_step_size=0.2, relaxation=1.5)
    out["sweeps__harmonic_relax15__opencv__f64__10"] = res
    # AMLE with 1-D convolutions (scipy.ndimage.convolve1d, symmetric border; amle_inpainter.py:26-37)
    res, _, _, _, _, _ = run_reference(
        classes["amle"], img, mask, convolver="opencv", max_iters=10, term_thres=1e-300,
        term_check_iters=10 ** 9, update_step_size=0.01, convolve_in_1d=True)
    out["sweeps__amle_1d__f64__10"] = res
    np.savez_compressed(os.path.join(HERE, "sweeps.npz"), **out)

    # ---------------------------------------------------------------- C. converged runs
    rows, cols = 64, 80
    mask = edge_mask(rows, cols, 0.30, 21)
    out = {"mask": mask}
    meta = []
    for dt_name, dtype in (("f64", np.float64), ("f32", np.float32)):
        img = surface(rows, cols, 9, dtype)
        img[~mask] = 0
        out["image__" + dt_name] = img
    cases = []
    for method in ("harmonic", "ccst", "tv", "amle"):
        for conv in ("opencv", "masked"):
            cases.append((method, conv, "f64", "relative", 1e-5, 10, {}))
    cases.append(("harmonic", "opencv", "f32", "relative", 1e-5, 10, {}))
    cases.append(("ccst", "opencv", "f32", "relative", 1e-4, 10, {}))
    cases.append(("harmonic", "opencv", "f64", "relative", 1e-8, 10, {}))
    cases.append(("harmonic", "opencv", "f64", "relative", 1e-5, 1000, {}))   # default cadence: 1001 sweeps
    cases.append(("harmonic", "opencv", "f64", "relative", 1e-5, 7, {}))
    cases.append(("harmonic", "opencv", "f64", "absolute", 1e-3, 10, {}))
    cases.append(("harmonic", "opencv", "f64", "absolute_percent", 1e-5, 10, {}))
    cases.append(("ccst", "masked", "f64", "absolute_percent", 1e-5, 25, {}))
    cases.append(("tv", "opencv", "f32", "absolute_percent", 1e-4, 10, {}))
    cases.append(("harmonic", "opencv", "f64", "relative", 1e-5, 10, {"init_with": "mean"}))
    cases.append(("harmonic", "opencv", "f64", "relative", 1e-9, 10, {"max_iters": 25}))  # not converged
    # fp32 runs of the one-sided schemes (appended: the keys above keep their index)
    cases.append(("amle", "opencv", "f32", "relative", 1e-4, 10, {}))
    cases.append(("amle", "masked", "f32", "relative", 1e-4, 10, {}))
    cases.append(("tv", "masked", "f32", "relative", 1e-4, 10, {}))
    cases.append(("tv", "masked", "f32", "absolute_percent", 1e-3, 10, {"max_iters": 4000}))
    cases.append(("amle", "masked", "f32", "absolute_percent", 1e-4, 10, {"max_iters": 4000}))
    for idx, (method, conv, dt_name, crit, thres, T, extra) in enumerate(cases):
        opts = dict(METHOD_OPTS[method])
        opts.update(convolver=conv, term_criteria=crit, term_thres=thres, term_check_iters=T)
        opts.update(extra)
        res, upd, d, term, img_after, inp = run_reference(
            classes[method], out["image__" + dt_name], mask, **opts)
        key = "conv%02d" % idx
        out[key] = res
        meta.append("|".join([key, method, conv, dt_name, crit, repr(thres), str(T),
                              repr(sorted(extra.items())), str(upd), repr(d), str(int(term)),
                              repr(float(inp.term_thres))]))
        print(meta[-1])
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "converged.npz"), **out)

    # ---------------------------------------------------------------- D. progress table (stdout)
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        run_reference(classes["harmonic"], out["image__f64"], mask, convolver="opencv",
                      update_step_size=0.2, term_criteria="relative", term_thres=1e-5,
                      term_check_iters=10, print_progress=True, print_progress_iters=7)
        run_reference(classes["harmonic"], out["image__f64"], mask, convolver="opencv",
                      update_step_size=0.2, term_criteria="relative", term_thres=1e-9,
                      term_check_iters=10, print_progress=True, print_progress_iters=4,
                      max_iters=23)
    with open(os.path.join(HERE, "progress_stdout.txt"), "w") as fh:
        fh.write(buf.getvalue())

    # ---------------------------------------------------------------- E. multigrid driver
    rows, cols = 101, 131                      # odd sizes: ceil-halving and non-2x up-sampling
    mask = edge_mask(rows, cols, 0.35, 31)
    mask[40:62, 50:75] = False                 # a hole that needs the coarse levels
    out = {"mask": mask}
    meta = []
    mg_cases = [("harmonic", "opencv", "f64", "relative", 1e-6, 3, 20, "zeros"),
                ("ccst", "masked", "f64", "absolute_percent", 1e-6, 2, 20, "mean"),
                ("harmonic", "opencv", "f32", "relative", 1e-5, 3, 20, "zeros"),
                ("tv", "opencv", "f64", "relative", 1e-5, 4, 30, "zeros"),      # pyramid stops early
                ("harmonic", "opencv", "f64", "relative", 1e-6, 2, 200, "zeros")]   # too small: single grid, NO init
    for dt_name, dtype in (("f64", np.float64), ("f32", np.float32)):
        img = surface(rows, cols, 13, dtype)
        img[~mask] = 0
        out["image__" + dt_name] = img
    for idx, (method, conv, dt_name, crit, thres, levels, min_res, init) in enumerate(mg_cases):
        opts = dict(METHOD_OPTS[method])
        opts.update(convolver=conv, term_criteria=crit, term_thres=thres, term_check_iters=10,
                    mgs_levels=levels, mgs_min_res=min_res, init_with=init, max_iters=20000)
        inp = classes[method](**opts)
        per_level = []
        orig_grid = inp.inpaint_grid
        orig_check = inp.term_check
        state = {"n": 0}

        def counting(f, fnew, _o=orig_check, _s=state):
            _s["n"] += 1
            return _o(f, fnew)

        def grid(image, mask_, _o=orig_grid, _s=state, _pl=per_level):
            _s["n"] = 0
            r = _o(image, mask_)
            _pl.append((image.shape[0], image.shape[1], (_s["n"] - 1) * 10 + 1))
            return r
        inp.term_check = counting
        inp.inpaint_grid = grid
        caller = out["image__" + dt_name].copy()
        with contextlib.redirect_stdout(io.StringIO()):
            res = inp.inpaint(caller, mask)
        key = "mg%02d" % idx
        out[key] = res
        out[key + "_caller_after"] = caller
        meta.append("|".join([key, method, conv, dt_name, crit, repr(thres), str(levels), str(min_res), init,
                              repr(per_level), repr(float(inp.term_thres))]))
        print(meta[-1])
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "multigrid.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
