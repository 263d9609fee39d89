"""CPU-side tests (no GPU): the C-ABI library loads and exports what include/hmi_b200.h declares,
the reference-facing Python layer keeps the reference's options / defaults / errors, and the
event-driven host loop reproduces the reference's check cadence and progress table."""
import argparse
import contextlib
import ctypes
import io
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from helpers import METHOD_OPTS
from oracle.fdpde_oracle import OracleInpainter

import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import _capi as C
from heightmap_interpolation_b200.apps.apps_common import create_inpainter_from_params
from heightmap_interpolation_b200.solver import Norm, make_params


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hmi_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hmi_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(C.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(C.SYMBOLS), declared ^ set(C.SYMBOLS)
    assert C.lib().hmi_abi_version() == 1


def test_params_struct_layout_matches_header():
    # 14 int32 + 5 double + 2 int64
    assert ctypes.sizeof(C.HmiParams) == 14 * 4 + 5 * 8 + 2 * 8
    p = make_params(8, 8, np.float64, C.HMI_HARMONIC)
    assert p.struct_size == ctypes.sizeof(C.HmiParams)


def test_create_rejects_bad_arguments_before_touching_the_gpu():
    lib = C.lib()
    h = ctypes.c_void_p()
    p = make_params(8, 8, np.float64, C.HMI_HARMONIC, dt=0.2)
    p.struct_size = 4
    assert lib.hmi_create(ctypes.byref(p), ctypes.byref(h)) == C.HMI_ERR_BAD_ARG
    for bad in (dict(rows=1), dict(dt=0.0), dict(term_thres=-1.0), dict(max_iters=0),
                dict(relaxation=0.5), dict(relaxation=2.5), dict(term_check_iters=0)):
        kw = dict(rows=8, cols=8, dtype=np.float64, method=C.HMI_HARMONIC, dt=0.2)
        kw.update(bad)
        p = make_params(**kw)
        assert lib.hmi_create(ctypes.byref(p), ctypes.byref(h)) == C.HMI_ERR_BAD_ARG, bad
        assert lib.hmi_last_error()
    p = make_params(8, 8, np.float64, C.HMI_CCST, dt=0.01, tension=1.5)
    assert lib.hmi_create(ctypes.byref(p), ctypes.byref(h)) == C.HMI_ERR_BAD_ARG
    assert b"tension" in lib.hmi_last_error()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    inp = hmi.SobolevInpainter(update_step_size=0.2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inp.inpaint(np.zeros((8, 8)), np.ones((8, 8), dtype=bool))


def test_nearest_and_multigrid_paths_fail_loudly_without_a_device():
    """The new device-side paths (nearest initialiser, fused fill, multigrid chaining) have no CPU
    fallback either; bad arguments are rejected before any CUDA call."""
    import ctypes
    import torch
    lib = C.lib()
    img, mask = np.zeros((8, 8)), np.ones((8, 8), dtype=np.uint8)
    assert lib.hmi_init_nearest_host(None, 64, mask.ctypes.data, 8, 8, 8, C.HMI_F64, 0) == C.HMI_ERR_BAD_ARG
    assert lib.hmi_init_nearest_host(img.ctypes.data, 8, mask.ctypes.data, 8, 8, 8, C.HMI_F64, 0) == C.HMI_ERR_BAD_ARG
    assert b"pitch" in lib.hmi_last_error()
    assert lib.hmi_set_term_thres(None, 1.0) == C.HMI_ERR_BAD_ARG
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hmi.SobolevInpainter(init_with="nearest").inpaint(np.zeros((8, 8)), np.eye(8, dtype=bool))
    big = np.zeros((1500, 1500))
    with pytest.raises(RuntimeError, match="no CPU fallback"):       # fused zeros/mean path (>= 2M cells)
        hmi.SobolevInpainter(init_with="mean").inpaint(big, np.ones(big.shape, dtype=bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hmi.SobolevInpainter(mgs_levels=2, mgs_min_res=20).inpaint(np.zeros((64, 64)), np.eye(64, dtype=bool))


def test_constructor_defaults_and_get_config_match_reference():
    inp = hmi.SobolevInpainter()
    assert inp.get_config() == {
        "update_step_size": 0.01, "term_criteria": "relative", "term_check_iters": 1000,
        "term_thres": 1e-8, "max_iters": int(1e8), "relaxation": 0, "print_progress": False,
        "print_progress_iters": 1000, "mgs_levels": 1, "mgs_min_res": 100, "init_with": "zeros",
        "convolver": "masked"}
    assert hmi.CCSTInpainter(tension=0.3).get_config()["tension"] == 0.3
    assert hmi.CCSTInpainter().get_config()["tension"] == 0.0
    assert hmi.TVInpainter().get_config()["epsilon"] == 1e-2
    assert "tension" not in hmi.AMLEInpainter().get_config()
    # unknown keywords are silently ignored, like the reference
    hmi.SobolevInpainter(backend="gpu", ti_arch="cuda", bogus=1)
    # floats are accepted for max_iters and truncated
    assert hmi.SobolevInpainter(max_iters=1e5).max_iters == 100000


@pytest.mark.parametrize("cls,kw,msg", [
    (hmi.SobolevInpainter, dict(update_step_size=0), "update_step_size must be larger than zero"),
    (hmi.SobolevInpainter, dict(term_thres=0), "term_thres must be larger than zero"),
    (hmi.SobolevInpainter, dict(max_iters=0), "max_iters must be larger than zero"),
    (hmi.SobolevInpainter, dict(relaxation=0.5), "relaxation must be a number between 1 and 2"),
    (hmi.SobolevInpainter, dict(relaxation=2.1), "relaxation must be a number between 1 and 2"),
    (hmi.SobolevInpainter, dict(mgs_levels=2.0), "mgs_levels must be an integer"),
    (hmi.SobolevInpainter, dict(mgs_min_res=10.5), "mgs_min_res must be an integer"),
    (hmi.SobolevInpainter, dict(init_with="harmonic"), "init_with must be one of these options"),
    (hmi.SobolevInpainter, dict(convolver="numpy"), "Unknown convolver type 'numpy'"),
    (hmi.CCSTInpainter, dict(tension=-0.1), "tension parameter must be a number between 0 and 1"),
    (hmi.CCSTInpainter, dict(tension=1.1), "tension parameter must be a number between 0 and 1"),
    (hmi.TVInpainter, dict(epsilon=0), "Epsilon parameter must be greater than zero"),
])
def test_constructor_validation_errors(cls, kw, msg):
    with pytest.raises(ValueError, match=msg):
        cls(**kw)


def test_unknown_term_criteria_raises_at_solve_time():
    inp = hmi.SobolevInpainter(term_criteria="bogus")       # accepted by the constructor
    with pytest.raises(ValueError, match="Unknown termination criteria!"):
        inp.set_term_criteria(np.zeros((4, 4)))


def test_absolute_percent_scales_threshold_in_place_and_compounds():
    inp = hmi.SobolevInpainter(term_criteria="absolute_percent", term_thres=1e-5)
    f = np.array([[0.0, 100.0], [50.0, 25.0]])
    inp.set_term_criteria(f)
    assert inp.term_thres == pytest.approx(1e-3)
    inp.set_term_criteria(f)                                   # the reference compounds on re-use
    assert inp.term_thres == pytest.approx(1e-1)


def test_convolver_names_map_to_orientation_and_border():
    tab = {n: (hmi.SobolevInpainter(convolver=n).convolver.orientation,
               hmi.SobolevInpainter(convolver=n).convolver.border)
           for n in ("opencv", "masked", "masked-parallel", "scipy-signal", "scipy-ndimage")}
    assert tab["opencv"] == (1, C.HMI_BORDER_REFLECT101)
    assert tab["masked"] == tab["masked-parallel"] == (-1, C.HMI_BORDER_REFLECT101)
    assert tab["scipy-signal"] == tab["scipy-ndimage"] == (-1, C.HMI_BORDER_ZERO)


def test_cli_dispatch_names():
    def ns(name, **kw):
        base = dict(subparser_name=name, update_step_size=0.2, term_thres=1e-5, verbose=False)
        base.update(kw)
        return argparse.Namespace(**base)
    assert isinstance(create_inpainter_from_params(ns("harmonic")), hmi.SobolevInpainter)
    tv = create_inpainter_from_params(ns("tv", epsilon=1.0))
    assert isinstance(tv, hmi.TVInpainter) and tv.epsilon == 1.0
    cc = create_inpainter_from_params(ns("ccst", tension=0.3))
    assert isinstance(cc, hmi.CCSTInpainter) and cc.tension == 0.3
    assert isinstance(create_inpainter_from_params(ns("ccst-ti", tension=0.1)), hmi.CCSTInpainter)
    assert isinstance(create_inpainter_from_params(ns("amle")), hmi.AMLEInpainter)
    # CLI defaults of the reference (apps_common.py:38-53)
    h = create_inpainter_from_params(ns("harmonic"))
    assert (h.term_criteria, h.init_with, h.convolver_type, h.max_iters) == \
        ("absolute_percent", "nearest", "opencv", 1000000)
    with pytest.raises(ValueError):
        create_inpainter_from_params(ns("telea"))


def test_initializers_zeros_and_mean_in_place():
    img = np.arange(12, dtype=np.float64).reshape(3, 4)
    mask = np.ones((3, 4), dtype=bool)
    mask[1, 1] = mask[2, 3] = False
    a = img.copy()
    out = hmi.SobolevInpainter(init_with="zeros").initializer.initialize(a, mask)
    assert out is a and a[1, 1] == 0 and a[2, 3] == 0
    b = img.copy()
    hmi.SobolevInpainter(init_with="mean").initializer.initialize(b, mask)
    assert b[1, 1] == pytest.approx(img[mask].mean())


class OracleEngine:
    """Stand-in for the CUDA solver in host-logic tests: same sweeps()/sweeps_norm() surface,
    arithmetic by the CPU oracle (test infrastructure)."""

    def __init__(self, method, image, mask, **opts):
        self.f = image.copy()
        self.mask = mask
        self.o = lambda k: OracleInpainter(method, max_iters=k, term_thres=1e-300,
                                           term_check_iters=10 ** 9, **opts)
        self.calls = []

    def sweeps(self, n):
        self.calls.append(("sweeps", n))
        if n > 0:
            self.f = self.o(n).inpaint_grid(self.f, self.mask)

    def sweeps_norm(self, n):
        self.calls.append(("norm", n))
        if n > 1:
            self.f = self.o(n - 1).inpaint_grid(self.f, self.mask)
        fnew = self.o(1).inpaint_grid(self.f, self.mask)
        d = fnew - self.f
        nm = Norm(float(np.sum(d * d)), float(np.sum(fnew * fnew)), float(np.abs(d).max()), 0.0)
        self.f = fnew
        return nm


def test_host_loop_reproduces_reference_progress_table_and_stop_iteration(golden_converged):
    mask = golden_converged["mask"]
    img = golden_converged["image__f64"].copy()
    img[~mask] = 0
    expected = open(os.path.join(GOLDEN, "progress_stdout.txt")).read()

    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        for thres, P, extra in ((1e-5, 7, {}), (1e-9, 4, {"max_iters": 23})):
            inp = hmi.SobolevInpainter(update_step_size=0.2, convolver="opencv", term_thres=thres,
                                       term_check_iters=10, print_progress=True,
                                       print_progress_iters=P, **extra)
            inp.print_msg("*** INPAINTING ***")
            inp.print_msg("* Initializing the inpainting problem using the 'zeros' filler")
            inp.print_msg("* Optimization:")
            inp.set_term_criteria(img)
            eng = OracleEngine("harmonic", img, mask, update_step_size=0.2, convolver="opencv")
            done, diff, conv = inp._loop(eng, C.HMI_RELATIVE)
            if not conv:
                print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")
            if not extra:
                assert (done, conv) == (51, True)
                assert np.allclose(eng.f, golden_converged["conv00"], rtol=0, atol=1e-9)
            else:
                assert (done, conv) == (23, False)
    assert buf.getvalue() == expected


def test_host_loop_check_cadence():
    """Checks fall on i % T == 0; returns after i+1 updates; no check when max_iters comes first."""
    class Fake:
        def __init__(self, converge_at):
            self.n = 0
            self.converge_at = converge_at
            self.log = []

        def sweeps(self, n):
            self.n += n
            self.log.append(("s", n))

        def sweeps_norm(self, n):
            self.n += n
            self.log.append(("n", n))
            return Norm(0.0 if self.n >= self.converge_at else 1.0, 1.0, 0.0, 0.0)

    inp = hmi.SobolevInpainter(term_check_iters=1000, term_thres=1e-5)
    f = Fake(converge_at=500)
    assert inp._loop(f, C.HMI_RELATIVE) == (1001, 0.0, True)
    assert f.log == [("n", 1), ("n", 1000)]
    inp = hmi.SobolevInpainter(term_check_iters=7, term_thres=1e-5, max_iters=20)
    f = Fake(converge_at=10 ** 9)
    done, diff, conv = inp._loop(f, C.HMI_RELATIVE)
    assert (done, conv) == (20, False)
    assert f.log == [("n", 1), ("n", 7), ("n", 7), ("s", 5)]
    # NaN never terminates (0/0 in the relative criterion)
    assert not (Norm(0.0, 0.0, 0.0, 0.0).diff(C.HMI_RELATIVE) < 1.0)
    assert np.isnan(Norm(0.0, 1.0, 3.0, 1.0).diff(C.HMI_ABSOLUTE))


def test_bilinear_resize_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    from heightmap_interpolation_b200.inpainting.multigrid import pyramid_shapes, resize_bilinear
    rng = np.random.default_rng(0)
    for dt, tol in ((np.float64, 1e-13), (np.float32, 1e-6)):
        for sh, dh in (((50, 64), (100, 128)), ((51, 66), (101, 131)), ((26, 33), (51, 66)), ((37, 53), (74, 105))):
            src = (rng.standard_normal(sh) * 100).astype(dt)
            want = cv2.resize(src, dsize=(dh[1], dh[0]))
            got = resize_bilinear(src, dh[0], dh[1])
            assert got.dtype == want.dtype
            assert np.abs(got - want).max() <= tol * 300
    # pyramid geometry: ceil-halving per level (= every 2**l-th cell), stopping before a level below min_res
    assert pyramid_shapes((5, 7), 3, 1) == ([(5, 7), (3, 4), (2, 2)], None)
    assert pyramid_shapes((401, 300), 4, 100) == ([(401, 300), (201, 150)], 2)
    assert pyramid_shapes((64, 64), 1, 100) == ([(64, 64)], None)
    a = np.arange(35.0).reshape(5, 7)
    assert a[::4, ::4].shape == pyramid_shapes((5, 7), 3, 1)[0][2]


def test_inpaint_work_areas_crop_and_paste():
    """Loop body of the reference CLI (apps/interpolate_netcdf4.py:176-212) with a stand-in inpainter:
    bounding-box crop, NaN -> 0 in the crop only, cells outside the area or not flagged stay untouched."""
    from heightmap_interpolation_b200.apps.interpolate import inpaint_work_areas
    elev = np.arange(30 * 40, dtype=np.float64).reshape(30, 40)
    mask_int = np.zeros((30, 40), dtype=bool)
    mask_int[5:25, 5:35] = True
    elev[mask_int] = np.nan
    area = np.zeros((30, 40, 1), dtype=bool)
    area[8:20, 10:30, 0] = True
    seen = {}

    class Fake:
        def inpaint(self, image, mask):
            seen["shape"], seen["nan"] = image.shape, bool(np.isnan(image).any())
            seen["known"] = int(mask.sum())
            out = image.copy()
            out[~mask] = 42.0
            return out

    got = inpaint_work_areas(elev, mask_int, area, inpainter_factory=Fake)
    assert seen["shape"] == (12, 20) and not seen["nan"] and seen["known"] == 0
    assert np.all(got[8:20, 10:30] == 42.0)
    outside = mask_int.copy()
    outside[8:20, 10:30] = False
    assert np.isnan(got[outside]).all()                       # holes outside the work area stay holes
    assert np.array_equal(got[~mask_int], elev[~mask_int])
    assert np.isnan(elev[8:20, 10:30]).all()                  # the input is not modified
    with pytest.raises(ValueError):
        inpaint_work_areas(elev, mask_int[:10], area, inpainter_factory=Fake)


def test_host_known_minmax_matches_numpy():
    """hmi_host_known_minmax (z-range of set_term_criteria, fd_pde_inpainter.py:149-155): min / max over
    the known cells, NaN-propagating like np.max / np.min, threaded above 65536 cells."""
    import ctypes
    rng = np.random.default_rng(0)
    for n, dt, code in ((1000, np.float64, C.HMI_F64), (300000, np.float32, C.HMI_F32)):
        img = rng.standard_normal(n).astype(dt)
        mask = rng.random(n) > 0.4
        lo, hi, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        m8 = mask.view(np.uint8)
        C.check(C.lib().hmi_host_known_minmax(img.ctypes.data, m8.ctypes.data, n, code, ctypes.byref(lo),
                                              ctypes.byref(hi), ctypes.byref(cnt)))
        assert (lo.value, hi.value, cnt.value) == (float(img[mask].min()), float(img[mask].max()), int(mask.sum()))
        img[np.nonzero(mask)[0][n // 3]] = np.nan
        C.check(C.lib().hmi_host_known_minmax(img.ctypes.data, m8.ctypes.data, n, code, ctypes.byref(lo),
                                              ctypes.byref(hi), ctypes.byref(cnt)))
        assert np.isnan(lo.value) and np.isnan(hi.value)
        img[:] = np.nan
        img[mask] = 1.0                                    # NaNs only in unknown cells do not count
        C.check(C.lib().hmi_host_known_minmax(img.ctypes.data, m8.ctypes.data, n, code, ctypes.byref(lo),
                                              ctypes.byref(hi), ctypes.byref(cnt)))
        assert (lo.value, hi.value) == (1.0, 1.0)


def test_slab_rows_needed_covers_ghosts():
    from heightmap_interpolation_b200.distributed import slab_bounds, slab_rows_needed
    rows, world = 1000, 4
    for rank in range(world):
        a, b, E = slab_rows_needed(rows, world, rank, radius=2, exchange_every=12)
        r0, r1 = slab_bounds(rows, world)[rank]
        assert E == 12
        assert a == (r0 - 24 if rank > 0 else 0) and b == (r1 + 24 if rank < world - 1 else rows)
    # the exchange depth is clamped so that the ghost rows fit inside the neighbouring slab
    assert slab_rows_needed(40, 4, 1, radius=2, exchange_every=12)[2] == 4


# ------------------------------------------------------------------ copies overlapped with the solve: launch order
def _band_steps(nb, m, phase):
    import ctypes
    cap = 4 * (nb + m + 2) * (nb + 2)
    buf = (ctypes.c_int32 * (4 * cap))()
    n = C.check(C.lib().hmi_debug_band_steps(nb, m, phase, buf, cap))
    return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]


def _check_band_order(nb, m, steps, all_landed):
    """Replay a launch order on a model of the two ping-pong grids.  level[j] = launches band j has completed, so
    its grid (level & 1) holds time level `level` and the other grid still holds time level level - 1.  Launch q of
    band j reads time level q of bands j-1, j, j+1 (halo rows) and overwrites what band j kept of time level q - 1."""
    level = [0] * nb
    landed = [all_landed] * nb
    finals = []
    for q, lo, hi, final in steps:
        if q < 0:                                   # band lo has landed in HBM
            assert not landed[lo] and (lo == 0 or landed[lo - 1])
            landed[lo] = True
            continue
        assert 0 <= lo <= hi < nb and 0 <= q < m
        for j in range(lo, hi + 1):
            assert landed[j] and level[j] == q, (nb, m, q, j, level[j])
        for nbr in (lo - 1, hi + 1):                # the bands next to the range provide halo rows at time level q
            if 0 <= nbr < nb:
                assert landed[nbr], (nb, m, q, nbr)
                # q: it is at that level; q + 1: it has moved on, but time level q is still in its other grid
                assert q <= level[nbr] <= q + 1, (nb, m, q, lo, hi, nbr, level[nbr])
        for j in range(lo, hi + 1):
            level[j] = q + 1
        if final >= 0:
            assert level[final] == m
            finals.append(final)
    assert level == [m] * nb and all(landed)
    return finals


def test_overlapped_copies_launch_order_respects_every_dependence():
    """csrc/hmi_capi.cu band_head_start_steps / band_drain_steps, the orders upload_pipelined and drain_pipelined
    queue on one stream: no launch reads a time level that is not there yet or not there any more."""
    for nb in list(range(1, 20)) + [31, 32, 33, 64]:
        steps = _band_steps(nb, nb, 0)
        assert _check_band_order(nb, nb, steps, all_landed=False) == []
        # under the upload every launch is one band; the catch-up is nb whole-range launches
        assert sum(1 for q, lo, hi, _ in steps if q >= 0 and hi > lo) == max(nb - 1, 0)
        for m in sorted({1, 2, 3, nb // 2, nb - 1, nb} - {0}):
            finals = _check_band_order(nb, m, _band_steps(nb, m, 1), all_landed=True)
            assert finals == list(range(nb)), (nb, m, finals)      # bands become final top to bottom, each once


# ------------------------------------------------------------------ wave-aware chunk picker
def _pick(nrows, nwin, halo, min_chunk=48, n_edge=2, edge_cost=1.7, slots=444, wpb=4):
    import ctypes
    a, b = ctypes.c_int32(), ctypes.c_int32()
    C.check(C.lib().hmi_debug_pick_chunks(nrows, nwin, n_edge, halo, wpb, slots, edge_cost, min_chunk,
                                          ctypes.byref(a), ctypes.byref(b)))
    rounds = -(-((n_edge * -(-nrows // b.value) + (nwin - n_edge) * -(-nrows // a.value) + wpb - 1) // wpb) // slots)
    return a.value, b.value, rounds


def test_chunk_picker_reproduces_the_measured_optima():
    """profiles/r05_tune_picker_min_rounds.log: the number of rounds of resident CTAs (444 on a B200) that measured
    fastest when it was forced, for the launches the load-balance term was fitted on — and the launches it must leave
    alone."""
    assert _pick(4096, 293, 8)[2] == 4               # 4096 x 32768 fp32 slab, harmonic K = 8 / CCST K = 4 (halo 8)
    assert _pick(512, 631, 8)[2] == 3                # a 512-row band of the overlapped copies, CCST fp64 K = 3
    assert _pick(4096, 79, 8)[2] == 2                # 4096^2 CCST fp64
    assert _pick(4096, 631, 8)[2] == 5               # 4096 x 32768 fp64 slab: unchanged
    ch, che, rounds = _pick(32768, 631, 8)           # the bench workload: many rounds, chunks of a few hundred rows
    assert 28 <= rounds <= 36 and 300 <= ch <= 450 and che < ch
    # fp32 TV / AMLE (edge strips 3.5x, chunks of at least 96 rows): 4096^2 stays one round of ~100-row chunks
    ch, che, rounds = _pick(4096, 35, 2, min_chunk=96, edge_cost=3.5)
    assert rounds == 1 and 96 <= ch <= 110
    # tiny launches: one chunk per strip, never taller than the launch
    assert _pick(20, 3, 6)[0] <= 20 and _pick(7, 1, 2, n_edge=1)[0] <= 8
