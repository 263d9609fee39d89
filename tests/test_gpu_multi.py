"""Row slabs over several GPUs: the peer-memory halo links (csrc/hmi_halo.cu), the one-process
device group behind ``Inpainter(..., devices=[...])`` (csrc/hmi_group.cu) and the one-process-per-GPU
path of distributed.py (CUDA IPC halo links + NCCL all-reduce; NCCL send/recv as the baseline).

Bar: a slab solve is BIT-IDENTICAL to the single-GPU solve of the whole grid and stops at the same
iteration (same kernels, same per-cell arithmetic, mirror-invariant neighbour sums), and therefore
inherits its parity with the reference.  Slabs may share a GPU, so the halo protocol itself is
covered on a one-GPU box too; the multi-process tests need two GPUs.
"""
import contextlib
import io
import os
import socket
import sys

import numpy as np
import pytest

from helpers import METHOD_OPTS, range_err
from oracle.fdpde_oracle import OracleInpainter

import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import _capi as C
from heightmap_interpolation_b200 import synthetic
from heightmap_interpolation_b200.solver import Group, Solver, make_params

pytestmark = pytest.mark.gpu

METHODS = {"harmonic": C.HMI_HARMONIC, "ccst": C.HMI_CCST, "tv": C.HMI_TV, "amle": C.HMI_AMLE}
CLASSES = {"harmonic": hmi.SobolevInpainter, "ccst": hmi.CCSTInpainter, "tv": hmi.TVInpainter,
           "amle": hmi.AMLEInpainter}


def n_gpus():
    return C.lib().hmi_device_count()


def device_lists():
    """Slab -> device assignments to test: everything on GPU 0, and (if present) real peers."""
    out = [[0, 0], [0, 0, 0]]
    n = n_gpus()
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


def problem(rows, cols, dtype, seed=3):
    img = synthetic.bathymetry((rows, cols), dtype=dtype)
    mask = synthetic.mask_random((rows, cols), 0.4, seed=seed)
    return synthetic.zero_unknown(img, mask), mask


def params_for(method, rows, cols, dtype, **kw):
    o = METHOD_OPTS[method]
    return make_params(rows, cols, dtype, METHODS[method], orientation=1, dt=o["update_step_size"],
                       tension=o.get("tension", 0.0), epsilon=o.get("epsilon", 1e-2), **kw)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("method", list(METHODS))
def test_group_sweeps_bit_identical_to_one_gpu(method, dtype):
    rows, cols = 700 + 37, 515
    img, mask = problem(rows, cols, dtype)
    K = 50          # several exchanges (every 12 / 18 sweeps) and a ragged last block
    with Solver(params_for(method, rows, cols, dtype)) as s:
        s.set_problem_host(img, mask)
        s.sweeps(K)
        nm1 = s.sweeps_norm(7)
        want = s.get_result_host()
    for devs in device_lists():
        with Group(params_for(method, rows, cols, dtype), devs) as g:
            assert g.n_slabs == len(devs) and g.ghost_rows > 0
            g.set_problem_host(img, mask)
            g.sweeps(K)
            nm = g.sweeps_norm(7)
            got = g.get_result_host()
        assert np.array_equal(got, want), (method, dtype, devs, range_err(got, want))
        assert np.array_equal(got[mask], img[mask])
        assert nm.max_abs == nm1.max_abs
        assert abs(nm.sum_d2 - nm1.sum_d2) <= 1e-12 * nm1.sum_d2 and abs(nm.sum_f2 - nm1.sum_f2) <= 1e-12 * nm1.sum_f2


def test_group_small_exchange_depth_and_tiny_slabs():
    """exchange_every = 1 (an exchange behind every sweep) and slabs only a few rows high: the split of
    the last temporal block has nothing to overlap with and must still hand over fresh rows."""
    rows, cols, dtype = 96, 130, np.float64
    img, mask = problem(rows, cols, dtype, seed=5)
    for method, E in (("ccst", 1), ("harmonic", 2), ("tv", 3), ("ccst", 5)):
        with Solver(params_for(method, rows, cols, dtype)) as s:
            s.set_problem_host(img, mask)
            s.sweeps(23)
            want = s.get_result_host()
        with Group(params_for(method, rows, cols, dtype), [0, 0, 0, 0], exchange_every=E) as g:
            assert g.exchange_every == E
            g.set_problem_host(img, mask)
            g.sweeps(23)
            got = g.get_result_host()
        assert np.array_equal(got, want), (method, E, range_err(got, want))


@pytest.mark.parametrize("method,dtype", [("harmonic", np.float64), ("ccst", np.float64), ("ccst", np.float32),
                                          ("tv", np.float32), ("amle", np.float64)])
def test_devices_keyword_converged_solve_equals_one_gpu_and_oracle(method, dtype):
    """The drop-in call with devices=[...]: same stop iteration and bits as one GPU; fp64 also against the
    CPU oracle at the 1e-10 bar (relative and absolute_percent criteria, zeros and mean initialisers)."""
    rows, cols = 640, 500
    img, mask = problem(rows, cols, dtype)
    for crit, thres, init in (("relative", 1e-4, "zeros"), ("absolute_percent", 1e-4, "mean")):
        opts = dict(convolver="opencv", term_criteria=crit, term_thres=thres, term_check_iters=25, max_iters=150,
                    init_with=init, **METHOD_OPTS[method])
        one = CLASSES[method](**opts)
        with contextlib.redirect_stdout(io.StringIO()):
            want = one.inpaint(img.copy(), mask)
        for devs in device_lists():
            inp = CLASSES[method](devices=devs, **opts)
            mine = img.copy()
            with contextlib.redirect_stdout(io.StringIO()):
                got = inp.inpaint(mine, mask)
            assert inp.iterations_ == one.iterations_ and inp.converged_ == one.converged_, (method, crit, devs)
            assert np.array_equal(got, want), (method, dtype, crit, devs, range_err(got, want))
            assert inp.term_thres == one.term_thres
        if dtype == np.float64 and method != "amle":
            # (AMLE normalises the gradient; on this zero-initialised bathymetry with kilometre-high jumps the
            # 1e-10 bar against the oracle is held by test_gpu_parity on smooth data, not here)
            ora = OracleInpainter(method, **opts)
            ref = ora.inpaint(img.copy(), mask)
            assert ora.updates_done == one.iterations_
            assert range_err(want, ref) <= 1e-10


def test_devices_keyword_large_grid_fused_init_and_progress_path():
    """Above the fused-initialiser size the one-shot C call drives the group (upload threads per device,
    the caller's array initialised in place after the upload); with print_progress the Python loop does."""
    rows, cols = 1600, 1400            # > 2M cells
    img, mask = problem(rows, cols, np.float32)
    raw = synthetic.bathymetry((rows, cols), dtype=np.float32)       # unknown cells still hold data
    opts = dict(convolver="opencv", term_criteria="absolute_percent", term_thres=1e-4, term_check_iters=20,
                max_iters=60, init_with="mean", update_step_size=0.01, tension=0.3)
    one = hmi.CCSTInpainter(**opts)
    a = raw.copy()
    with contextlib.redirect_stdout(io.StringIO()):
        want = one.inpaint(a, mask)
    for devs in device_lists()[:3]:
        inp = hmi.CCSTInpainter(devices=devs, **opts)
        b = raw.copy()
        with contextlib.redirect_stdout(io.StringIO()):
            got = inp.inpaint(b, mask)
        assert inp.iterations_ == one.iterations_
        assert np.array_equal(got, want)
        assert np.array_equal(a, b)                      # the caller's array was initialised in place, identically
        assert np.all(b[~mask] == b[~mask][0])
    buf = io.StringIO()
    inp = hmi.CCSTInpainter(devices=[0, 0], print_progress=True, print_progress_iters=20, **opts)
    with contextlib.redirect_stdout(buf):
        got = inp.inpaint(raw.copy(), mask)
    assert np.array_equal(got, want) and inp.iterations_ == one.iterations_
    assert "| Iteration | Function change |" in buf.getvalue()


def test_group_refuses_what_it_cannot_do():
    p = params_for("ccst", 256, 256, np.float64, ghost_top=4)
    with pytest.raises(ValueError):
        Group(p, [0, 0])
    with pytest.raises(ValueError):
        hmi.CCSTInpainter(devices=[]).inpaint(np.zeros((64, 64)), np.ones((64, 64), dtype=bool))
    with pytest.raises(ValueError):
        hmi.CCSTInpainter(devices="some").inpaint(np.zeros((64, 64)), np.ones((64, 64), dtype=bool))
    # a grid too small for the requested number of slabs falls back to fewer
    with Group(params_for("ccst", 40, 64, np.float64), [0, 0, 0, 0]) as g:
        assert g.n_slabs < 4


def test_unlinked_slab_refuses_exchange():
    p = params_for("harmonic", 128, 128, np.float64, ghost_top=6)
    with Solver(p) as s:
        s.set_problem_host(np.zeros((134, 128)), np.ones((134, 128), dtype=bool))
        with pytest.raises(C.HmiError):
            s.sweeps_exchange(3)


# ------------------------------------------------------------------ one process per GPU (torch.distributed)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, halo, cases, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from heightmap_interpolation_b200.distributed import inpaint_slab
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        for key, method, dtype, E in cases:
            rows, cols = 1024 + 37, 777
            img, mask = problem(rows, cols, dtype)
            opts = dict(convolver="opencv", term_criteria="relative", term_thres=1e-4, term_check_iters=25, max_iters=120,
                        device=rank, **METHOD_OPTS[method])
            inp = CLASSES[method](**opts)
            with contextlib.redirect_stdout(io.StringIO()):
                out, (r0, r1) = inpaint_slab(inp, img, mask, exchange_every=E, halo=halo)
            ret[(key, rank)] = (r0, r1, out, inp.iterations_, inp.exchanges_)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("halo", ["peer", "nccl"])
def test_one_process_per_gpu_slabs_bit_identical_to_one_gpu(halo):
    """Two ranks (torch.multiprocessing, NCCL) run inpaint_slab with the overlapped halo exchange: ghost
    rows pushed through CUDA IPC mappings (halo='peer') or sent with NCCL (halo='nccl'); the
    convergence sums are all-reduced with NCCL either way."""
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    cases = [("h64", "harmonic", np.float64, 8), ("c64", "ccst", np.float64, 6), ("c32", "ccst", np.float32, 12),
             ("t32", "tv", np.float32, 4), ("a64", "amle", np.float64, 2)]
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_rank_main, args=(world, _free_port(), halo, cases, ret), nprocs=world, join=True)
        ret = dict(ret)
    for key, method, dtype, E in cases:
        rows, cols = 1024 + 37, 777
        img, mask = problem(rows, cols, dtype)
        opts = dict(convolver="opencv", term_criteria="relative", term_thres=1e-4, term_check_iters=25, max_iters=120,
                    **METHOD_OPTS[method])
        one = CLASSES[method](**opts)
        with contextlib.redirect_stdout(io.StringIO()):
            want = one.inpaint(img.copy(), mask)
        parts = sorted((ret[(key, r)] for r in range(world)), key=lambda t: t[0])
        got = np.vstack([p[2] for p in parts])
        assert all(p[3] == one.iterations_ for p in parts), (key, [p[3] for p in parts], one.iterations_)
        assert parts[0][4] > 0
        assert np.array_equal(got, want), (key, halo, range_err(got, want))


# ------------------------------------------------------------------ multigrid with the pyramid on the device(s)
@pytest.mark.parametrize("method,dt,crit,init,levels,nans", [("harmonic", np.float64, "absolute_percent", "mean", 3, False),
                                                             ("ccst", np.float32, "relative", "zeros", 3, False),
                                                             ("tv", np.float64, "relative", "nearest", 3, False),
                                                             ("harmonic", np.float32, "relative", "mean", 2, True),
                                                             ("ccst", np.float64, "absolute_percent", "zeros", 2, True)])
def test_multigrid_on_device_groups_equals_one_gpu_and_oracle(method, dt, crit, init, levels, nans):
    """Every pyramid level is a device group: decimation from one upload, up-sampling of the coarser
    solution across slabs (peer reads), blend, z-range — bit-identical to the single-GPU run, same
    per-level iteration counts as the oracle's restatement of inpaint_multigrid (fp64: 1e-10).  With NaNs
    in the unknown cells (two levels: an intermediate level would be poisoned by NaN * 0, in the reference
    too) the coarsest initialiser writes through the strided view and level 0 scrubs the rest (:259-260)."""
    rows, cols = 333, 470                     # odd sizes at every level
    mask = synthetic.mask_holes((rows, cols), 12, 25, seed=4) & synthetic.mask_random((rows, cols), 0.2, seed=5)
    img = synthetic.bathymetry((rows, cols), dtype=dt)
    img[~mask] = np.nan if nans else 0
    opts = dict(METHOD_OPTS[method], convolver="opencv", term_criteria=crit, term_thres=1e-5 if dt == np.float64 else 1e-4,
                term_check_iters=20, max_iters=3000, mgs_levels=levels, mgs_min_res=50, init_with=init)
    one = CLASSES[method](**opts)
    a = img.copy()
    with contextlib.redirect_stdout(io.StringIO()):
        want = one.inpaint(a, mask)
    assert len(one.history_) == levels and not np.isnan(want).any()
    assert not np.isnan(a).any() and np.array_equal(a[mask], img[mask])      # scrubbed / initialised in place
    for devs in device_lists():
        inp = CLASSES[method](devices=devs, **opts)
        b = img.copy()
        with contextlib.redirect_stdout(io.StringIO()):
            got = inp.inpaint(b, mask)
        assert inp.history_ == one.history_, (devs, inp.history_, one.history_)
        assert np.array_equal(got, want), (method, devs, range_err(got, want))
        assert np.array_equal(a, b)
        assert float(inp.term_thres) == float(one.term_thres)
    if init != "nearest" and not nans:
        ora = OracleInpainter(method, **opts)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ora.inpaint(img.copy(), mask)
        assert one.history_ == ora.history
        assert range_err(want, ref) <= (1e-10 if dt == np.float64 else 1e-5)


# ------------------------------------------------------------------ the CLI's loop over work areas, on the device
def _cli_case(dt):
    rows, cols = 420, 610
    elev = synthetic.bathymetry((rows, cols), dtype=dt)
    mask_int = ~(synthetic.mask_holes((rows, cols), 10, 22, seed=8) & synthetic.mask_random((rows, cols), 0.25, seed=9))
    elev[mask_int] = np.nan                                   # what a NetCDF grid with holes looks like
    areas = np.zeros((rows, cols, 3), dtype=bool)
    areas[30:260, 40:400, 0] = True
    yy, xx = np.ogrid[:rows, :cols]
    areas[:, :, 1] = (yy - 250) ** 2 + (xx - 380) ** 2 < 150 ** 2          # overlaps area 0, not a rectangle
    areas[300:410, 5:120, 2] = True
    return elev, mask_int, areas


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("method,init,crit", [("harmonic", "zeros", "absolute_percent"), ("ccst", "mean", "relative"),
                                              ("tv", "zeros", "relative")])
def test_work_areas_on_the_device_equal_the_host_route(method, init, crit, dt):
    """inpaint_work_areas (apps/interpolate_netcdf4.py:176-212): crop, known mask, NaN -> 0, initialiser, solve
    and paste on the device against the same loop through host arrays — several overlapping areas, NaN
    holes, one GPU and device groups."""
    import argparse
    from heightmap_interpolation_b200.apps.interpolate import inpaint_work_areas
    elev, mask_int, areas = _cli_case(dt)
    opts = dict(METHOD_OPTS[method], convolver="opencv", term_criteria=crit, term_thres=1e-4, term_check_iters=20,
                max_iters=200, init_with=init)
    seen = []

    def factory(devices=None):
        def make():
            seen.append(CLASSES[method](devices=devices, **opts))
            return seen[-1]
        return make
    with contextlib.redirect_stdout(io.StringIO()):
        host = inpaint_work_areas(elev, mask_int, areas, inpainter_factory=factory(), on_device=False)
        host_iters = [s.iterations_ for s in seen if s.iterations_ is not None]
        for devs in [None] + device_lists()[:2]:
            del seen[:]
            dev = inpaint_work_areas(elev, mask_int, areas, inpainter_factory=factory(devs), on_device=True)
            assert [s.iterations_ for s in seen] == host_iters
            if init == "zeros":
                assert np.array_equal(dev, host, equal_nan=True), (method, devs)
            else:                       # the mean is accumulated in double on the device
                assert np.array_equal(np.isnan(dev), np.isnan(host))
                ok = ~np.isnan(host)
                assert np.abs(dev[ok] - host[ok]).max() <= (1e-9 if dt == np.float64 else 1e-2)
    inside = areas.any(axis=2) & mask_int
    assert not np.isnan(host[inside]).any()                               # every flagged cell inside an area was filled
    assert np.isnan(host[mask_int & ~areas.any(axis=2)]).all()            # holes outside the areas stay holes
    assert np.array_equal(host[~mask_int], elev[~mask_int])               # known cells untouched
    # no areas = one area covering the grid; the CLI namespace route
    ns = argparse.Namespace(subparser_name="harmonic", update_step_size=0.2, term_thres=1e-4, term_criteria="relative",
                            term_check_iters=20, max_iters=100, relaxation=0, mgs_levels=1, mgs_min_res=100, init_with="zeros",
                            convolver="opencv", debug_dir="", verbose=False, print_progress_iters=1000, backend="b200")
    with contextlib.redirect_stdout(io.StringIO()):
        a = inpaint_work_areas(elev, mask_int, None, ns)
        b = inpaint_work_areas(elev, mask_int, None, ns, on_device=False)
    assert np.array_equal(a, b) and not np.isnan(a).any()


def test_work_areas_device_route_is_refused_when_the_initialiser_needs_the_host():
    from heightmap_interpolation_b200.apps.interpolate import inpaint_work_areas
    elev, mask_int, areas = _cli_case(np.float64)
    with pytest.raises(ValueError):
        inpaint_work_areas(elev, mask_int, areas, inpainter_factory=lambda: hmi.SobolevInpainter(init_with="nearest"),
                           on_device=True)
