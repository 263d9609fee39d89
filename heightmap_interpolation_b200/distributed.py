"""Row-slab domain decomposition across the GPUs of one box (one process per GPU).

The reference is a single process holding the whole grid; this module is the only
parallelism the path admits: every sweep is a local stencil, so the grid is cut into contiguous row
slabs, each slab carries `ghost` rows of its neighbours above and below, and

  * every `exchange_every` sweeps the ghost rows are refreshed.  On the GPUs this is the peer-memory
    halo link of csrc/hmi_halo.cu (`halo="peer"`, the default): the slab's copy engine pushes its
    rows next to a seam into a staging area in the neighbour's HBM (mapped through CUDA IPC, the
    64-byte handles are all-gathered once at set-up) and publishes a flag there; no SM and no
    collective are involved, and the push overlaps the interior update of the same temporal block.
    `halo="nccl"` keeps the grouped send/recv per neighbour of round 1 as the comparison baseline;
    the CPU tests (gloo) use the same send/recv path.  Between exchanges the solver advances a
    shrinking trapezoid of ghost rows itself (see hmi_sweeps in include/hmi_b200.h), so
    ghost = radius * exchange_every rows buy `exchange_every` sweeps per exchange;
  * on a check iteration the per-slab convergence sums are all-reduced (sum, sum, max, max; NCCL) —
    the only global quantity of the loop (fd_pde_inpainter.py:368-379);
  * the reflect-101 / zero border rule applies only at the global top and bottom, never at seams.

`torch.distributed` is plumbing only: it moves 64-byte handles and four doubles (and, with
`halo="nccl"`, the ghost rows).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi as C
from .inpainting.fd_pde_inpainter import _RangeView
from .solver import Norm, Solver


def slab_bounds(rows, world):
    """Contiguous, near-equal row ranges [r0, r1) for each rank."""
    base, rem = divmod(int(rows), int(world))
    out, r0 = [], 0
    for k in range(world):
        r1 = r0 + base + (1 if k < rem else 0)
        out.append((r0, r1))
        r0 = r1
    return out


class _DevArray:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2,
                                         "strides": strides}


class CudaSlabEngine:
    """The CUDA solver of one slab plus zero-copy torch views of its ghost/edge rows."""

    def __init__(self, params, halo="peer"):
        if halo not in ("peer", "nccl"):
            raise ValueError("halo must be 'peer' or 'nccl'")
        self.solver = Solver(params)
        self.params = params
        self.device = torch.device("cuda", params.device)
        self.typestr = "<f8" if params.dtype == C.HMI_F64 else "<f4"
        self.elem = 8 if params.dtype == C.HMI_F64 else 4
        self._views = {}            # device pointer -> torch view (the solver ping-pongs two grids)
        self._side = None           # side stream of the overlapped exchange
        self.halo = halo
        self.linked = False
        self._group = None

    def link(self, rank, world, group=None):
        """Connect this slab to its neighbours' halo blocks (CUDA IPC): one all-gather of the 64-byte
        handles, then cudaIpcOpenMemHandle of the rank above and below."""
        if self.halo != "peer" or world == 1 or (self.params.ghost_top == 0 and self.params.ghost_bottom == 0):
            return
        info = self.solver.halo_setup()
        mine = torch.tensor(list(bytes(info.ipc_handle)), dtype=torch.uint8, device=self.device)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=group)
        if rank > 0:
            self.solver.halo_connect_ipc(0, bytes(handles[rank - 1].cpu().tolist()))
        if rank < world - 1:
            self.solver.halo_connect_ipc(1, bytes(handles[rank + 1].cpu().tolist()))
        dist.barrier(group=group)      # every block is mapped before anybody pushes
        self.linked = True
        self._group = group

    def sweeps_exchange(self, n):
        """n sweeps, then the push of the rows next to the seams into the neighbours' HBM, overlapped
        with the interior update (hmi_sweeps_exchange)."""
        main = torch.cuda.current_stream(self.device)
        self.solver.sweeps_exchange(n, main.cuda_stream, 0)

    def load(self, image, mask):
        self.solver.set_problem_host(image, mask)

    def _grid(self):
        ptr, pitch = self.solver.current_grid()
        t = self._views.get(ptr)
        if t is None:
            rows = self.solver.phys_rows
            arr = _DevArray(ptr, (rows, pitch // self.elem), self.typestr)
            t = self._views[ptr] = torch.as_tensor(arr, device=self.device)
        return t

    def rows_view(self, lo, hi):
        """Physical rows [lo, hi) of the current iterate (row 0 = first ghost row), pitched."""
        return self._grid()[lo:hi]

    def sweeps(self, n):
        self.solver.sweeps(n)

    def sweeps_overlap(self, n, exchange):
        """n sweeps with the halo exchange of the *result* overlapped with the interior update: the
        edge rows are produced first on a side stream, `exchange()` is queued behind them on that
        stream, and the main stream only waits for it before the next block."""
        main = torch.cuda.current_stream(self.device)
        if self._side is None:     # high priority: the exchange should not queue behind the interior update
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
        self.solver.sweeps_overlap(n, main.cuda_stream, self._side.cuda_stream)
        with torch.cuda.stream(self._side):
            exchange()
        main.wait_stream(self._side)

    def sweeps_norm(self, n):
        return self.solver.sweeps_norm(n)

    def result(self):
        return self.solver.get_result_host()

    def close(self):
        if self.linked:                # nobody may still be pushing into a block about to be freed
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self._group)
            self.linked = False
        self.solver.close()


class SlabSolver:
    """One rank's share of a row-slab decomposed solve.  Provides the sweeps()/sweeps_norm()
    surface FDPDEInpainter._loop drives, so the reference's loop semantics are shared with the
    single-GPU path."""

    def __init__(self, engine, rank, world, radius, ghost_top, ghost_bottom, exchange_every,
                 rows, group=None):
        self.engine, self.rank, self.world, self.group = engine, rank, world, group
        self.radius = radius
        self.gt, self.gb = ghost_top, ghost_bottom
        self.rows = rows
        self.exchange_every = int(exchange_every)
        if world > 1 and self.exchange_every * radius > max(ghost_top, ghost_bottom):
            raise ValueError("exchange_every*radius exceeds the ghost depth")
        self.exchanges = 0
        self.halo_fresh = True      # slabs are loaded with their neighbours' rows
        self.overlap = hasattr(engine, "sweeps_overlap")
        # peer-memory halo links: the library tracks the freshness of the ghost rows itself
        self.peer = bool(getattr(engine, "linked", False))

    # -- halo exchange: my first/last `g` interior rows -> the neighbour's ghost rows
    def exchange(self):
        if self.world == 1:
            return
        ops, keep = [], []
        e, gt, gb, R = self.engine, self.gt, self.gb, self.rows
        if self.rank > 0:                      # neighbour above
            send = e.rows_view(gt, gt + gt)    # my top gt interior rows (its bottom ghost depth == gt)
            recv = e.rows_view(0, gt)
            ops += [dist.P2POp(dist.isend, send, self.rank - 1, self.group),
                    dist.P2POp(dist.irecv, recv, self.rank - 1, self.group)]
            keep += [send, recv]
        if self.rank < self.world - 1:         # neighbour below
            send = e.rows_view(gt + R - gb, gt + R)
            recv = e.rows_view(gt + R, gt + R + gb)
            ops += [dist.P2POp(dist.isend, send, self.rank + 1, self.group),
                    dist.P2POp(dist.irecv, recv, self.rank + 1, self.group)]
            keep += [send, recv]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.exchanges += 1

    def _blocks(self, n):
        E = self.exchange_every if self.world > 1 else max(n, 1)
        while n > 0:
            k = min(n, E)
            yield k
            n -= k

    def _block(self, k):
        """k <= exchange_every sweeps.  Halos are refreshed before the block if they are stale;
        engines that can split their last launch refresh them again *after* it, overlapped with
        the interior update, so the next block starts without waiting."""
        if self.peer:
            self.engine.sweeps_exchange(k)
            self.exchanges += 1
            return
        if not self.halo_fresh:
            self.exchange()
        if self.overlap and self.world > 1:
            self.engine.sweeps_overlap(k, self.exchange)
            self.halo_fresh = True
        else:
            self.engine.sweeps(k)
            self.halo_fresh = False

    def sweeps(self, n):
        for k in self._blocks(int(n)):
            self._block(k)

    def sweeps_norm(self, n):
        n = int(n)
        blocks = list(self._blocks(n))
        for k in blocks[:-1]:
            self._block(k)
        if not self.halo_fresh and not self.peer:
            self.exchange()
        nm = self.engine.sweeps_norm(blocks[-1])
        self.halo_fresh = False
        return self.allreduce_norm(nm)

    def allreduce_norm(self, nm):
        if self.world == 1:
            return nm
        dev = getattr(self.engine, "device", torch.device("cpu"))
        sums = torch.tensor([nm.sum_d2, nm.sum_f2], dtype=torch.float64, device=dev)
        maxs = torch.tensor([nm.max_abs, nm.has_nan], dtype=torch.float64, device=dev)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=self.group)
        sums, maxs = sums.cpu(), maxs.cpu()
        return Norm(float(sums[0]), float(sums[1]), float(maxs[0]), float(maxs[1]))


def slab_rows_needed(rows, world, rank, radius, exchange_every):
    """Rows [a, b) of the global grid rank `rank` must hold (its slab plus ghost rows) and the
    exchange depth actually used — what `inpaint_slab(..., first_row=a, global_rows=rows)` expects."""
    bounds = slab_bounds(rows, world)
    r0, r1 = bounds[rank]
    min_rows = min(b - a for a, b in bounds)
    E = max(1, min(int(exchange_every), (min_rows - 1) // radius if world > 1 else 10 ** 9))
    ghost = radius * E if world > 1 else 0
    return r0 - (ghost if rank > 0 else 0), r1 + (ghost if rank < world - 1 else 0), E


def inpaint_slab(inpainter, image, mask, rank=None, world=None, exchange_every=None,
                 engine_factory=None, group=None, global_rows=None, first_row=0, post_hook=None, halo="peer",
                 out=None):
    """Distributed `inpaint_grid`: every rank passes the same global `image`/`mask` (already
    initialised, host arrays), or — so that no process ever holds the whole grid — only the rows
    [first_row, first_row + len(image)) of a grid of `global_rows` rows, which must cover this
    rank's slab and ghost rows (`slab_rows_needed`); returns this rank's interior rows and its
    (r0, r1).  With partial arrays the z-range of `absolute_percent` is all-reduced.

    The loop semantics (check cadence, stop rule, progress table on rank 0) are those of
    FDPDEInpainter._loop.  `engine_factory(params)` exists for the CPU tests (gloo + a CPU engine);
    the product path always builds CudaSlabEngine.  `post_hook(slab)` (benchmarks) runs after the solve
    while the slab is still resident.
    """
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    rows, cols = image.shape
    partial = global_rows is not None
    if partial:
        rows = int(global_rows)
    radius = 2 if inpainter.METHOD == C.HMI_CCST else 1
    bounds = slab_bounds(rows, world)
    r0, r1 = bounds[rank]
    min_rows = min(b - a for a, b in bounds)
    if exchange_every is None:      # a whole number of temporal blocks (harmonic 6 / fp32 8, TV/AMLE 2, CCST 3 / fp32 4 sweeps per launch)
        harmonic_f32 = inpainter.METHOD == C.HMI_HARMONIC and image.dtype == np.float32
        exchange_every = (16 if harmonic_f32 else 18) if radius == 1 else 12
    exchange_every = max(1, min(int(exchange_every), (min_rows - 1) // radius if world > 1 else 10 ** 9))
    ghost = radius * exchange_every if world > 1 else 0
    gt = ghost if rank > 0 else 0
    gb = ghost if rank < world - 1 else 0

    device = inpainter._device() if engine_factory is None else 0
    if partial:
        a, b = r0 - gt - first_row, r1 + gb - first_row
        if a < 0 or b > image.shape[0]:
            raise ValueError("rows [%d, %d) needed, the arrays hold [%d, %d)"
                             % (r0 - gt, r1 + gb, first_row, first_row + image.shape[0]))
        image, mask = image[a:b], mask[a:b]
        _set_term_criteria_global(inpainter, image[gt:gt + (r1 - r0)], world, group,
                                  torch.device("cuda", device) if engine_factory is None else torch.device("cpu"))
        lo = 0
    else:
        inpainter.set_term_criteria(image)   # every rank sees the same global z-range
        lo = r0 - gt
    params = inpainter._make_params(r1 - r0, cols, image.dtype, ghost_top=gt, ghost_bottom=gb,
                                    device=device)
    engine = engine_factory(params) if engine_factory else CudaSlabEngine(params, halo=halo)
    try:
        engine.load(np.ascontiguousarray(image[lo:lo + gt + (r1 - r0) + gb]),
                    np.ascontiguousarray(mask[lo:lo + gt + (r1 - r0) + gb]))
        if hasattr(engine, "link"):
            engine.link(rank, world, group)
        slab = SlabSolver(engine, rank, world, radius, gt, gb, exchange_every, r1 - r0, group)
        quiet = inpainter.print_progress and rank != 0
        if quiet:
            inpainter.print_progress = False
        try:
            done, diff, conv = inpainter._loop(slab, params.criteria)
        finally:
            if quiet:
                inpainter.print_progress = True
        inpainter._record(done, diff, conv, getattr(engine, "kernel_name", None), None)
        inpainter.exchanges_ = slab.exchanges
        if not conv and rank == 0:
            print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")
        result = engine.result() if out is None else engine.solver.get_result_host(out)
        if post_hook is not None:
            post_hook(slab)
        return result, (r0, r1)
    finally:
        engine.close()


def _set_term_criteria_global(inpainter, interior, world, group, dev):
    """set_term_criteria (fd_pde_inpainter.py:141-158) when no rank holds the whole grid: the
    z-range of 'absolute_percent' is the all-reduced max/min of the slabs."""
    if inpainter.term_criteria != "absolute_percent" or world == 1:
        inpainter.set_term_criteria(interior)
        return
    t = torch.tensor([float(np.max(interior)), -float(np.min(interior))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    t = t.cpu()
    inpainter.set_term_criteria(_RangeView(interior.dtype.type(-float(t[1])), interior.dtype.type(float(t[0]))))
