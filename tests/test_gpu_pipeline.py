"""The one-shot call's pipelined copies (csrc/hmi_capi.cu: upload_pipelined, drain_pipelined): the grid goes up in
row bands and the bands that have landed run iteration 0 and their first temporal blocks while the rest is still
on the PCIe bus; a solve that ends at max_iters runs its last launches band-skewed and copies every band out as
soon as it is final.  Bar: BIT-IDENTICAL to upload-then-solve (same kernels, and they do not care how sweeps are cut into
launches or row ranges), same stop iteration; last_diff of iteration 0 may differ in its last bits (band sums
folded in band order).  HMI_UPLOAD_PIPELINE=2 forces the route on small grids, =0 turns it off.
"""
import contextlib
import io

import numpy as np
import pytest

from helpers import DTYPES, METHOD_OPTS

import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import _capi as C
from heightmap_interpolation_b200 import synthetic

pytestmark = pytest.mark.gpu

CLASSES = {"harmonic": hmi.SobolevInpainter, "ccst": hmi.CCSTInpainter, "tv": hmi.TVInpainter,
           "amle": hmi.AMLEInpainter}


def solve(monkeypatch, mode, method, img, mask, **opts):
    monkeypatch.setenv("HMI_UPLOAD_PIPELINE", str(mode))
    before = C.lib().hmi_upload_pipeline_count()
    inp = CLASSES[method](**dict(METHOD_OPTS[method], **opts))
    work = img.copy()
    with contextlib.redirect_stdout(io.StringIO()):
        out = inp.inpaint(work, mask)
    return out, work, inp, C.lib().hmi_upload_pipeline_count() - before


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("method", list(CLASSES))
def test_pipelined_upload_is_bit_identical(monkeypatch, method, dt):
    rows, cols = 1531, 1483          # > 2M cells: the fused zeros / mean initialiser route as well
    mask = synthetic.mask_random((rows, cols), 0.4, seed=11)
    img = synthetic.bathymetry((rows, cols), dtype=DTYPES[dt])
    img[~mask] = np.nan if method == "harmonic" else 0       # unknown cells are never read before the initialiser
    cases = [dict(init_with="zeros", term_check_iters=1000, max_iters=61, term_thres=1e-300),        # no check after i = 0
             dict(init_with="mean", term_check_iters=25, max_iters=391, term_thres=1e-300),          # checks, never converges, short drain
             dict(init_with="zeros", term_check_iters=30, max_iters=121, term_thres=1e-300),          # the last request carries a check
             dict(init_with="zeros", term_check_iters=30, max_iters=5000,                            # shallow pipeline, converges
                  term_thres=3e-4 if method in ("harmonic", "ccst") else 3e-3)]
    for opts in cases:
        want, work0, a, n0 = solve(monkeypatch, 0, method, img, mask, convolver="opencv", **opts)
        got, work2, b, n2 = solve(monkeypatch, 2, method, img, mask, convolver="opencv", **opts)
        assert n0 == 0 and n2 == 1, (n0, n2)
        assert b.iterations_ == a.iterations_ and b.converged_ == a.converged_, (method, dt, opts)
        assert np.array_equal(got, want), (method, dt, opts)
        assert np.array_equal(work2, work0)                  # the caller's array is initialised in place either way
        if a.last_diff_ == a.last_diff_:
            assert abs(b.last_diff_ - a.last_diff_) <= 1e-9 * abs(a.last_diff_)


def test_pipelined_upload_masked_orientation_and_converged_at_iteration_zero(monkeypatch):
    rows, cols = 900, 1000
    img = synthetic.bathymetry((rows, cols))
    # every cell known: iteration 0 changes nothing, the reference returns after one update and so must we,
    # although the bands have run ahead (the call falls back to the plain route)
    mask = np.ones((rows, cols), dtype=bool)
    got, _, inp, n = solve(monkeypatch, 2, "ccst", img, mask, term_check_iters=50, max_iters=500, term_thres=1e-8)
    assert n == 1 and inp.iterations_ == 1 and inp.converged_
    assert np.array_equal(got, img)
    # convolution orientation (the class default 'masked' convolver), TV with three checks
    mask = synthetic.mask_random((rows, cols), 0.3, seed=2)
    img[~mask] = 0
    opts = dict(convolver="masked", term_check_iters=30, max_iters=95, term_thres=1e-300)
    want, _, a, _ = solve(monkeypatch, 0, "tv", img, mask, **opts)
    got, _, b, n = solve(monkeypatch, 2, "tv", img, mask, **opts)
    assert n == 1 and b.iterations_ == a.iterations_ == 95
    assert np.array_equal(got, want)


def test_pipelined_copies_through_inpaint_grid_and_a_preallocated_result(monkeypatch):
    """inpaint_grid() on an already initialised array (hmi_inpaint_host: no fused initialiser) into out=."""
    rows, cols = 1100, 1300
    mask = synthetic.mask_random((rows, cols), 0.35, seed=4)
    img = synthetic.bathymetry((rows, cols), dtype=np.float32)
    img[~mask] = -2500.0
    opts = dict(METHOD_OPTS["amle"], convolver="opencv", term_check_iters=40, max_iters=130, term_thres=1e-300)
    res = {}
    for mode in (0, 2):
        monkeypatch.setenv("HMI_UPLOAD_PIPELINE", str(mode))
        before = C.lib().hmi_upload_pipeline_count()
        out = np.full((rows, cols), np.nan, dtype=np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            got = hmi.AMLEInpainter(**opts).inpaint_grid(img, mask, out=out)
        assert got is out and C.lib().hmi_upload_pipeline_count() - before == (1 if mode else 0)
        res[mode] = out
    assert np.array_equal(res[2], res[0]) and not np.isnan(res[2]).any()


def test_pipelined_upload_is_the_default_on_large_grids_and_not_on_small_ones(monkeypatch):
    monkeypatch.delenv("HMI_UPLOAD_PIPELINE", raising=False)
    lib = C.lib()
    n = 8192                                                  # 2^26 cells, 16 bands of 512 rows
    mask = synthetic.mask_random(n, 0.4, seed=3)
    img = synthetic.bathymetry(n, dtype=np.float32)
    img[~mask] = 0
    opts = dict(METHOD_OPTS["ccst"], convolver="opencv", term_check_iters=100, max_iters=101, term_thres=1e-300)
    before = lib.hmi_upload_pipeline_count()
    with contextlib.redirect_stdout(io.StringIO()):
        got = hmi.CCSTInpainter(**opts).inpaint(img.copy(), mask)
    assert lib.hmi_upload_pipeline_count() == before + 1
    monkeypatch.setenv("HMI_UPLOAD_PIPELINE", "0")
    with contextlib.redirect_stdout(io.StringIO()):
        want = hmi.CCSTInpainter(**opts).inpaint(img.copy(), mask)
    assert lib.hmi_upload_pipeline_count() == before + 1
    assert np.array_equal(got, want)
    monkeypatch.delenv("HMI_UPLOAD_PIPELINE", raising=False)
    with contextlib.redirect_stdout(io.StringIO()):
        hmi.CCSTInpainter(**opts).inpaint(img[:2048, :2048].copy(), mask[:2048, :2048])
    assert lib.hmi_upload_pipeline_count() == before + 1     # 4M cells: plain route
