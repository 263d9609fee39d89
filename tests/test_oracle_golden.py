"""The CPU oracle (oracle/) against fixtures produced by the unmodified reference.

This is what pins the oracle: every later CUDA parity test compares
against the oracle, so the oracle itself must reproduce the reference's outputs first.
"""
import numpy as np
import pytest

from helpers import DTYPES, METHOD_OPTS, parse_converged_meta, range_err
from oracle.fdpde_oracle import OracleInpainter

METHODS = ["harmonic", "ccst", "tv", "amle"]
# f64: the oracle and the reference differ only by summation order inside a 3x3 sum.
ORACLE_TOL = {"f64": 1e-12, "f32": 2e-6}


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("conv", ["opencv", "masked", "scipy-ndimage", "scipy-signal"])
@pytest.mark.parametrize("method", METHODS)
def test_step_fun_matches_reference(golden_step, method, conv, dt):
    f = golden_step["field"].astype(DTYPES[dt])
    ref = golden_step["step__%s__%s__%s" % (method, conv, dt)]
    work = np.ones(f.shape, dtype=np.uint8) if conv.startswith("masked") else None
    got = OracleInpainter(method, convolver=conv, **METHOD_OPTS[method]).step_fun(f, work)
    if not conv.startswith("scipy"):   # scipy.signal.convolve promotes f32 x f64-kernel to f64
        assert got.dtype == ref.dtype
    assert range_err(got, ref) <= ORACLE_TOL[dt]


@pytest.mark.parametrize("K", [1, 10, 60])
@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("conv", ["opencv", "masked"])
@pytest.mark.parametrize("method", METHODS)
def test_k_sweeps_match_reference(golden_sweeps, method, conv, dt, K):
    img = golden_sweeps["image__" + dt].copy()
    mask = golden_sweeps["mask"]
    ref = golden_sweeps["sweeps__%s__%s__%s__%d" % (method, conv, dt, K)]
    o = OracleInpainter(method, convolver=conv, max_iters=K, term_thres=1e-300,
                        term_check_iters=10 ** 9, **METHOD_OPTS[method])
    got = o.inpaint(img, mask)
    assert o.updates_done == K and not o.converged
    assert np.array_equal(got[mask], ref[mask])          # known cells bit-exact
    assert range_err(got, ref) <= ORACLE_TOL[dt] * max(1, K // 10)


@pytest.mark.parametrize("conv", ["scipy-ndimage", "scipy-signal"])
@pytest.mark.parametrize("method", METHODS)
def test_other_border_rules(golden_sweeps, method, conv):
    img = golden_sweeps["image__f64"].copy()
    mask = golden_sweeps["mask"]
    ref = golden_sweeps["sweeps__%s__%s__f64__10" % (method, conv)]
    o = OracleInpainter(method, convolver=conv, max_iters=10, term_thres=1e-300,
                        term_check_iters=10 ** 9, **METHOD_OPTS[method])
    assert range_err(o.inpaint(img, mask), ref) <= 1e-12


@pytest.mark.parametrize("t", [0, 1])
def test_ccst_tension_end_points(golden_sweeps, t):
    img = golden_sweeps["image__f64"].copy()
    mask = golden_sweeps["mask"]
    ref = golden_sweeps["sweeps__ccst_t%d__opencv__f64__10" % t]
    o = OracleInpainter("ccst", convolver="opencv", max_iters=10, term_thres=1e-300,
                        term_check_iters=10 ** 9, update_step_size=0.01, tension=float(t))
    assert range_err(o.inpaint(img, mask), ref) <= 1e-12


def test_over_relaxation_f64(golden_sweeps):
    img = golden_sweeps["image__f64"].copy()
    mask = golden_sweeps["mask"]
    ref = golden_sweeps["sweeps__harmonic_relax15__opencv__f64__10"]
    o = OracleInpainter("harmonic", convolver="opencv", max_iters=10, term_thres=1e-300,
                        term_check_iters=10 ** 9, update_step_size=0.2, relaxation=1.5)
    got = o.inpaint(img, mask)
    assert np.array_equal(got[mask], ref[mask])
    assert range_err(got, ref) <= 1e-12


def test_converged_runs_match_reference(golden_converged):
    mask = golden_converged["mask"]
    for c in parse_converged_meta(golden_converged):
        img = golden_converged["image__" + c["dtype"]].copy()
        opts = dict(METHOD_OPTS[c["method"]])
        opts.update(convolver=c["convolver"], term_criteria=c["criteria"], term_thres=c["thres"],
                    term_check_iters=c["T"])
        opts.update(c["extra"])
        o = OracleInpainter(c["method"], **opts)
        got = o.inpaint(img, mask)
        ref = golden_converged[c["key"]]
        assert o.updates_done == c["updates"], c
        assert o.converged == c["converged"], c
        assert o.last_diff == pytest.approx(c["diff"], rel=1e-3 if c["dtype"] == "f32" else 1e-7), c
        assert o.term_thres == pytest.approx(c["final_thres"], rel=1e-12), c  # absolute_percent mutation
        assert np.array_equal(got[mask], ref[mask]), c
        assert range_err(got, ref) <= (1e-11 if c["dtype"] == "f64" else 5e-6), c


def test_amle_convolve_in_1d(golden_sweeps):
    img = golden_sweeps["image__f64"].copy()
    mask = golden_sweeps["mask"]
    ref = golden_sweeps["sweeps__amle_1d__f64__10"]
    o = OracleInpainter("amle", convolver="opencv", max_iters=10, term_thres=1e-300,
                        term_check_iters=10 ** 9, update_step_size=0.01, convolve_in_1d=True)
    assert range_err(o.inpaint(img, mask), ref) <= 1e-12


def test_nearest_initialiser_matches_reference(golden_nearest):
    """Initializer 'nearest' (initializer.py:19-34, scipy griddata): the restatement returns the
    reference's value wherever the nearest known cell is unique, and a known cell at exactly the
    reference's (minimal) distance elsewhere — SciPy's choice among equidistant cells is an
    artefact of its KD-tree leaf order."""
    from oracle.fdpde_oracle import nearest_init
    g = golden_nearest
    for row in g["meta"]:
        key = str(row).split("|")[0]
        img, mask = g[key + "_image"].copy(), g[key + "_mask"]
        got, d2 = nearest_init(img, mask, return_d2=True)
        ref, uniq = g[key + "_init"], g[key + "_unique"]
        assert got is img                                   # in place, like the reference
        assert np.array_equal(got[mask], ref[mask])
        assert np.array_equal(d2, g[key + "_d2"]), key
        assert np.array_equal(got[uniq], ref[uniq]), key
        assert int((~uniq).sum()) > 0                       # the fixture does exercise ties


def test_solve_from_reference_nearest_initialisation(golden_nearest):
    """From the reference's own nearest initialisation the oracle reproduces the reference solve
    (CLI-style options: absolute_percent, opencv) update for update."""
    g = golden_nearest
    upd, diff, term, final_thres = str(g["nn00_solve_meta"][0]).split("|")
    o = OracleInpainter("harmonic", update_step_size=0.2, term_criteria="absolute_percent", term_thres=1e-6,
                        term_check_iters=10, max_iters=100000, convolver="opencv")
    got = o.inpaint_grid(g["nn00_init"].copy(), g["nn00_mask"])
    assert (o.updates_done, o.converged) == (int(upd), bool(int(term)))
    assert o.term_thres == pytest.approx(float(final_thres), rel=1e-12)
    assert range_err(got, g["nn00_solve"]) <= 1e-12


def test_multigrid_driver_matches_reference(golden_multigrid):
    """The oracle's restatement of inpaint_multigrid (fd_pde_inpainter.py:187-279) against the
    reference fixtures: per-level grid sizes and update counts, the result, the write-through into
    the caller's array and the final (mutated) threshold."""
    import ast
    g = golden_multigrid
    mask = g["mask"]
    for row in g["meta"]:
        key, method, conv, dt, crit, thres, levels, min_res, init, per_level, final_thres = str(row).split("|")
        opts = dict(METHOD_OPTS[method])
        opts.update(convolver=conv, term_criteria=crit, term_thres=float(thres), term_check_iters=10,
                    mgs_levels=int(levels), mgs_min_res=int(min_res), init_with=init, max_iters=20000)
        o = OracleInpainter(method, **opts)
        caller = g["image__" + dt].copy()
        got = o.inpaint(caller, mask)
        assert o.history == ast.literal_eval(per_level), (key, o.history)
        ref = g[key]
        assert got.dtype == ref.dtype
        assert np.array_equal(got[mask], ref[mask]), key
        assert range_err(got, ref) <= (1e-11 if dt == "f64" else 5e-6), (key, range_err(got, ref))
        assert np.allclose(caller, g[key + "_caller_after"], rtol=0, atol=1e-9 if dt == "f64" else 1e-3), key
        assert float(o.term_thres) == pytest.approx(float(final_thres), rel=1e-6), key


def test_product_resize_equals_oracle_resize():
    """The host tables the multigrid kernel consumes (product) and the oracle's cv2.resize restatement
    are two independent statements of the same rule."""
    from heightmap_interpolation_b200.inpainting.multigrid import resize_bilinear as product_resize
    from oracle.fdpde_oracle import resize_bilinear as oracle_resize
    rng = np.random.default_rng(5)
    for dt in (np.float64, np.float32):
        for (sr, sc), (dr, dc) in (((51, 66), (101, 131)), ((26, 33), (51, 66)), ((40, 40), (80, 80)),
                                   ((7, 9), (13, 18))):
            src = rng.standard_normal((sr, sc)).astype(dt)
            assert np.array_equal(product_resize(src, dr, dc), oracle_resize(src, dr, dc))


def test_fp32_amle_is_ill_conditioned_on_jump_data_in_the_oracle_itself():
    """Pins the statement behind the one skipped GPU combination (test_mid_sized_odd_grids_streaming_equals_generic,
    AMLE fp32; DESIGN.md "AMLE in fp32"): on the 1501 x 2047 bathymetry grid with 40-cell holes the CPU
    restatement of the reference run in fp32 differs from the same restatement in fp64 by 7.5e-3 of the
    range after 7 sweeps, at cell (861, 1097) — the direction g/|g| of amle_inpainter.py:53-57 is decided
    by rounding noise there — while almost every other cell agrees to fp32 accuracy."""
    from heightmap_interpolation_b200 import synthetic
    rows, cols = 1501, 2047
    mask = synthetic.mask_holes((rows, cols), 30, 40, seed=3) & synthetic.mask_random((rows, cols), 0.3, seed=4)
    mask[0, :] = mask[-1, :] = False
    mask[:, 0] = mask[:, -1] = False
    res = {}
    for dt in ("f64", "f32"):
        img = synthetic.bathymetry((rows, cols), dtype=DTYPES[dt])
        img[~mask] = 0
        o = OracleInpainter("amle", convolver="opencv", update_step_size=0.01, max_iters=7, term_thres=1e-300,
                            term_check_iters=10 ** 9)
        res[dt] = o.inpaint_grid(img, mask).astype(np.float64)
    d = np.abs(res["f64"] - res["f32"]) / float(res["f64"].max() - res["f64"].min())
    worst = np.unravel_index(d.argmax(), d.shape)
    assert worst == (861, 1097)
    assert 5e-3 < d.max() < 1e-2                         # 7.52e-3: three orders above the fp32 bar of 1e-5
    assert (d > 1e-3).sum() <= 10 and (d > 1e-5).sum() <= 2000   # 5 and 1058 cells of 3.07 million
    assert np.median(d[~mask]) < 1e-7
