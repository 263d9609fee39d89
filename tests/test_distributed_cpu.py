"""World-size-2 (and 3) gloo tests of the row-slab host logic on CPU: partition, halo exchange,
trapezoid ghost advance between exchanges, convergence all-reduce, shared loop semantics."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import METHOD_OPTS
from oracle.fdpde_oracle import OracleInpainter

import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200 import synthetic
from heightmap_interpolation_b200.distributed import inpaint_slab, slab_bounds, slab_rows_needed

CLASSES = {"harmonic": hmi.SobolevInpainter, "ccst": hmi.CCSTInpainter, "tv": hmi.TVInpainter}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(rows=90, cols=70):
    img = synthetic.gaussian_hill((rows, cols)) + 3.0
    mask = synthetic.mask_random((rows, cols), 0.35, seed=4)
    mask[0, :5] = False
    mask[-1, -7:] = False
    return synthetic.zero_unknown(img, mask), mask


def _worker(rank, world, port, method, opts, exchange_every, q, partial=False):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleSlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img, mask = _problem()
        inp = CLASSES[method](**opts)
        if partial:      # the rank only ever holds its own rows (+ ghosts) of the grid
            radius = 2 if method == "ccst" else 1
            a, b, _ = slab_rows_needed(img.shape[0], world, rank, radius, exchange_every)
            out, (r0, r1) = inpaint_slab(inp, img[a:b].copy(), mask[a:b].copy(), exchange_every=exchange_every,
                                         engine_factory=OracleSlabEngine, global_rows=img.shape[0], first_row=a)
        else:
            out, (r0, r1) = inpaint_slab(inp, img, mask, exchange_every=exchange_every,
                                         engine_factory=OracleSlabEngine)
        q.put((rank, r0, r1, out, inp.iterations_, inp.last_diff_, inp.converged_, inp.exchanges_))
    finally:
        dist.destroy_process_group()


def _run(world, method, opts, exchange_every, partial=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, opts, exchange_every, q, partial))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


def test_slab_bounds_cover_rows():
    assert slab_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab_bounds(4096, 8)[-1] == (3584, 4096)


@pytest.mark.parametrize("world,method,exchange_every", [(2, "harmonic", 4), (2, "ccst", 3),
                                                         (3, "tv", 2), (2, "harmonic", 1)])
def test_slab_solve_equals_single_grid(world, method, exchange_every):
    opts = dict(METHOD_OPTS[method], convolver="opencv", term_thres=1e-5, term_check_iters=10,
                max_iters=400)
    img, mask = _problem()
    ref = OracleInpainter(method, **opts)
    want = ref.inpaint_grid(img.copy(), mask)
    res = _run(world, method, opts, exchange_every)
    got = np.vstack([r[3] for r in res])
    assert [(r[1], r[2]) for r in res] == slab_bounds(img.shape[0], world)
    for r in res:
        assert r[4] == ref.updates_done and r[6] == ref.converged
        assert r[5] == pytest.approx(ref.last_diff, rel=1e-9)
        # slabs are loaded with fresh halos, so the first block needs no exchange
        assert r[7] >= (ref.updates_done - 1) // max(exchange_every, 1)
    assert np.array_equal(got, want)       # same arithmetic per cell: bit-exact


def test_partial_slabs_absolute_percent():
    """Each rank holds only its slab (+ ghost rows); the z-range of 'absolute_percent' is all-reduced
    (set_term_criteria, fd_pde_inpainter.py:149-155) and the solve equals the single-grid one."""
    opts = dict(METHOD_OPTS["ccst"], convolver="opencv", term_criteria="absolute_percent", term_thres=1e-4,
                term_check_iters=10, max_iters=400)
    img, mask = _problem()
    ref = OracleInpainter("ccst", **opts)
    want = ref.inpaint_grid(img.copy(), mask)
    res = _run(3, "ccst", opts, 3, partial=True)
    got = np.vstack([r[3] for r in res])
    for r in res:
        assert r[4] == ref.updates_done and r[6] == ref.converged
    assert np.array_equal(got, want)
