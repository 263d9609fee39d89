#!/usr/bin/env python
"""Benchmark of the FD-PDE inpainting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): grid-cell updates per second, GLUPS = rows*cols*sweeps / s / 1e9.

Workload (BASELINE.json configs[4], SURVEY.md 8d "C5", the headline configuration and the largest one
that fits one GPU): CCST inpainting, tension 0.3, dt 0.01, of a 32768 x 32768 synthetic bathymetry grid
with 40 % of the cells missing (random mask, seed 3), fp64.  One *step* is SWEEPS_PER_STEP relaxation
sweeps of the whole grid (one default check window of the reference).  At N > 1 (torchrun, one rank per
GPU) the SAME grid is cut into N row slabs (strong scaling); ghost rows travel through peer-memory halo
links (csrc/hmi_halo.cu) every `exchange_every` sweeps and the convergence sums through an NCCL
all-reduce.

  value : whole-job GLUPS with the grid resident in HBM, timed with CUDA events on the launch stream
  e2e   : the same metric through the public API — Inpainter(**opts).inpaint(image, mask) at N = 1,
          distributed.inpaint_slab(inpainter, rows, mask, ...) at N > 1 — with pinned host buffers:
          H2D of image + mask and D2H of the result inside the timed region, every step.  At N = 1 the
          library overlaps both copies with the solve (row bands in anti-diagonal order, csrc/hmi_capi.cu);
          e2e.copies_overlapped counts the calls that did
  roofline.achieved : algorithmic bytes (17 B fp64 per grid-cell update) per launch /
          average launch duration; with temporal blocking (k sweeps per launch) this can exceed the HBM
          peak, so roofline.traffic carries the ncu-measured DRAM bytes per launch
  parity : checked inside the run — a short solve through the same public API against the CPU oracle on
          crops (corners with the global borders, slab seams), known cells bit-exact, and SHA-256 of the
          4096-row blocks of the e2e result against the hashes of the 1-GPU run (profiles/)
  cpu_baseline : the CPU oracle (oracle/, an OpenMP C port of the reference loop) on the host cores,
          on a bounded sample of the same workload (a band of the grid)

`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the reference is
pure Python + NumPy/OpenCV/Numba and cannot travel to the GPU box) on a bounded sample of the same
config, on rank 0 with all the host threads it can use.
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID = 32768               # C5: 32768 x 32768
MISSING, MASK_SEED = 0.40, 3
SWEEPS_PER_STEP = 1000     # one reference check window (term_check_iters default)
BLOCK = 4096               # rows per checksum block (a whole number per slab at N = 1, 2, 4, 8)
PARITY_SWEEPS = 36         # short solve checked against the oracle (three exchanges at N > 1)
BYTES_PER_LUP = {"f64": 17, "f32": 9}
CCST = dict(update_step_size=0.01, tension=0.3, convolver="opencv")
WORKLOAD = ("CCST tension=0.3 dt=0.01, %dx%d synthetic bathymetry, 40%% missing (random mask seed %d), fp64"
            % (GRID, GRID, MASK_SEED))
CPU_BAND_ROWS = 512        # the CPU legs run a band of the grid: rows [0, 512) x all 32768 columns
HASH_FILE = os.path.join(ROOT, "profiles", "c5_ccst_f64_1001_sweeps_block_sha256.json")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def generate_rows(a, b, dtype, img_out=None, mask_out=None, cols=(0, GRID)):
    """Rows [a, b) (columns [c0, c1)) of the C5 problem: bathymetry with the unknown cells zeroed (what the
    reference CLI does, apps/interpolate_netcdf4.py:192) and the mask (True = known)."""
    from heightmap_interpolation_b200 import synthetic
    c0, c1 = cols
    if img_out is None:
        img_out = np.empty((b - a, c1 - c0), dtype=dtype)
        mask_out = np.empty((b - a, c1 - c0), dtype=bool)
    for r in range(a, b, 1024):
        e = min(b, r + 1024)
        m = synthetic.mask_random((GRID, GRID), MISSING, seed=MASK_SEED, row_range=(r, e))[:, c0:c1]
        f = synthetic.bathymetry((GRID, GRID), seed=7, dtype=dtype, row_range=(r, e))[:, c0:c1]
        f[~m] = 0
        img_out[r - a:e - a] = f
        mask_out[r - a:e - a] = m
    return img_out, mask_out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------- CPU legs
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_oracle_all_cores():
    """The oracle with every host core: torchrun exports OMP_NUM_THREADS=1, so set it before libgomp
    is initialised and again through the OpenMP API."""
    import ctypes
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    from oracle import fdpde_oracle
    fdpde_oracle.lib()
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except Exception:
        pass
    return fdpde_oracle.OracleInpainter, n


def cpu_band():
    return generate_rows(0, CPU_BAND_ROWS, np.float64)


def cpu_run(Oracle, img, mask, sweeps):
    o = Oracle("ccst", max_iters=sweeps, term_thres=1e-300, term_check_iters=10 ** 9, **CCST)
    t = time.perf_counter()
    o.inpaint_grid(img, mask)
    return time.perf_counter() - t


def cpu_port_glups(budget_s=15.0):
    """Time the CPU oracle on a bounded sample of the workload: CCST sweeps of a 512-row band of the
    same grid (128 MiB per array, far beyond the caches) for about budget_s, in one call."""
    Oracle, cores = load_oracle_all_cores()
    img, mask = cpu_band()
    cpu_run(Oracle, img, mask, 2)           # warm-up (thread pool, first touch)
    t1 = cpu_run(Oracle, img, mask, 10) / 10.0
    k = int(max(100, min(5000, budget_s / max(t1, 1e-6))))
    t = cpu_run(Oracle, img, mask, k)
    return img.size * k / t / 1e9, k, t, cores


def run_reference(args):
    """Reference arm: the reference algorithm on the host CPU (oracle port), same config and metric, every
    step one oracle call of REF_SWEEPS sweeps over the band."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Oracle, cores = load_oracle_all_cores()
    img, mask = cpu_band()
    sweeps = 120
    for _ in range(max(args.warmup, 1)):
        cpu_run(Oracle, img, mask, sweeps)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_run(Oracle, img, mask, sweeps)
    glups = img.size * sweeps * args.steps / dt / 1e9
    sample = ("%d CCST sweeps per step (one oracle call) of rows [0, %d) x %d columns of the %dx%d fp64 grid, "
              "OpenMP C port of the reference loop on %d threads" % (sweeps, CPU_BAND_ROWS, GRID, GRID, GRID, cores))
    print(json.dumps({
        "impl": "reference", "metric": "grid-cell updates/s", "value": glups, "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sweeps_per_step": sweeps, "sample": sample},
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------- parity
def block_hashes(arr, first_row):
    """SHA-256 (first 16 hex digits) of each BLOCK-row block of `arr` (rows first_row ...)."""
    out = {}
    for r in range(0, arr.shape[0], BLOCK):
        out[str((first_row + r) // BLOCK)] = hashlib.sha256(np.ascontiguousarray(arr[r:r + BLOCK]).data).hexdigest()[:16]
    return out


def oracle_crop_check(result_rows, first_row, r_lo, r_hi, c_lo, c_hi):
    """Rows [r_lo, r_hi) x columns [c_lo, c_hi) of a PARITY_SWEEPS-sweep solve against the CPU oracle run
    on the crop enlarged by the domain of dependence (2 cells per CCST sweep); where the enlarged crop
    hits the grid edge, the edge is the global border and the oracle applies the border rule there."""
    from oracle.fdpde_oracle import OracleInpainter
    m = 2 * PARITY_SWEEPS + 2
    a, b = max(0, r_lo - m), min(GRID, r_hi + m)
    c, d = max(0, c_lo - m), min(GRID, c_hi + m)
    img, mask = generate_rows(a, b, np.float64, cols=(c, d))
    o = OracleInpainter("ccst", max_iters=PARITY_SWEEPS, term_thres=1e-300, term_check_iters=10 ** 9, **CCST)
    ref = o.inpaint_grid(img, mask)[r_lo - a:r_hi - a, c_lo - c:c_hi - c]
    got = result_rows[r_lo - first_row:r_hi - first_row, c_lo:c_hi]
    known = mask[r_lo - a:r_hi - a, c_lo - c:c_hi - c]
    src = img[r_lo - a:r_hi - a, c_lo - c:c_hi - c]
    err = float(np.abs(got - ref).max()) / 4725.0       # z-range bound of the synthetic bathymetry
    return err, bool(np.array_equal(got[known], src[known]))


def parity_crops(r0, r1, world, rank):
    """Crops this rank checks: the grid corners it owns and the rows on both sides of its upper seam."""
    h = 96
    crops = []
    if rank == 0:
        crops += [("top-left", 0, h, 0, 128), ("top-right", 0, h, GRID - 128, GRID)]
    if rank == world - 1:
        crops += [("bottom-left", GRID - h, GRID, 0, 128), ("bottom-right", GRID - h, GRID, GRID - 128, GRID)]
    if rank > 0:
        crops += [("seam-%d-below" % rank, r0, r0 + h, 12000, 12128)]
    if rank < world - 1:
        crops += [("seam-%d-above" % (rank + 1), r1 - h, r1, 20000, 20128)]
    if world == 1:          # the seams an 8-GPU run would have, plus the middle of the grid
        crops += [("row-%d" % r, r - h // 2, r + h // 2, 12000, 12128) for r in (4096, 16384, 28672)]
    return crops


# ---------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import heightmap_interpolation_b200 as hmi
    from heightmap_interpolation_b200 import _capi as C
    from heightmap_interpolation_b200.distributed import inpaint_slab, slab_rows_needed
    from heightmap_interpolation_b200.solver import Solver, make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dtype, dkey = np.float64, "f64"
    S, E = args.sweeps, args.exchange_every
    peak, peak_src = measured_peak()
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- this rank's rows of the problem, generated straight into pinned host memory
    a, b, E = slab_rows_needed(GRID, world, rank, 2, E)
    r0, r1 = (GRID * rank) // world, (GRID * (rank + 1)) // world
    pin_img = torch.empty((b - a, GRID), dtype=torch.float64, pin_memory=True)
    pin_mask = torch.empty((b - a, GRID), dtype=torch.bool, pin_memory=True)
    pin_out = torch.empty((r1 - r0, GRID), dtype=torch.float64, pin_memory=True)
    a_img, a_mask, a_out = pin_img.numpy(), pin_mask.numpy(), pin_out.numpy()
    generate_rows(a, b, dtype, a_img, a_mask)

    def make_inpainter(max_iters, T):
        return hmi.CCSTInpainter(term_check_iters=T, max_iters=max_iters, term_thres=1e-300, init_with="zeros",
                                 device=local, temporal_block=args.temporal_block, **CCST)

    def solve(inp, out):
        """The public call: the whole grid on one GPU, or this rank's slab of it."""
        with contextlib.redirect_stdout(io.StringIO()):      # "[WARNING] did NOT converge" by design
            if world == 1:
                return inp.inpaint(a_img, a_mask, out=out), None
            return inpaint_slab(inp, a_img, a_mask, exchange_every=E, global_rows=GRID, first_row=a, out=out,
                                halo=args.halo, post_hook=solve.hook)
    solve.hook = None

    # ------------------------------------------------------------------ device-resident throughput
    timing = {}

    def timed_steps(step, launches):
        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        barrier()
        timing["ms"] = max_over_ranks(ev0.elapsed_time(ev1))
        timing["launches"] = launches() - l0
        timing["clocks"] = sampler.stop() if rank == 0 else None

    if world == 1:
        p = make_params(GRID, GRID, dtype, C.HMI_CCST, orientation=1, dt=CCST["update_step_size"],
                        tension=CCST["tension"], device=local, temporal_block=args.temporal_block)
        with Solver(p) as solver:
            if args.vec:
                solver.set_option("vec", args.vec)
            if args.tma:
                solver.set_option("tma", args.tma)
            solver.set_problem_host(a_img, a_mask)
            timed_steps(lambda: solver.sweeps(S), lambda: solver.launch_count)
            kname, tb = solver.kernel_name, solver.temporal_block
    else:
        def hook(slab):      # runs inside inpaint_slab while the slab is resident and linked
            timed_steps(lambda: slab.sweeps(S), lambda: slab.engine.solver.launch_count)
            timing["kernel"] = (slab.engine.solver.kernel_name, slab.engine.solver.temporal_block)
        solve.hook = hook
        solve(make_inpainter(1, 1), a_out)
        solve.hook = None
        kname, tb = timing["kernel"]
    ms = timing["ms"]
    value = float(GRID) * GRID * S * steps / (ms * 1e-3) / 1e9

    # ------------------------------------------------------------------ end to end through the public API
    e2e_sweeps = S + 1          # one default check window: checks at sweeps 1 and S+1
    inp = None
    overlapped0 = C.lib().hmi_upload_pipeline_count()
    for _ in range(2):
        solve(make_inpainter(e2e_sweeps, S), a_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        inp = make_inpainter(e2e_sweeps, S)
        solve(inp, a_out)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / steps
    assert inp.iterations_ == e2e_sweeps
    overlapped = int(C.lib().hmi_upload_pipeline_count() - overlapped0) - 2      # (two untimed calls first)
    e2e_val = float(GRID) * GRID * e2e_sweeps / e2e_s / 1e9
    bytes_in = torch.tensor([a_img.nbytes + a_mask.nbytes, a_out.nbytes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(bytes_in)
    h2d, d2h = int(bytes_in[0].item()), int(bytes_in[1].item())

    # ------------------------------------------------------------------ parity, inside the run
    hashes = block_hashes(a_out, r0)                         # of the e2e result (S + 1 sweeps)
    par_out = np.empty_like(a_out)
    solve(make_inpainter(PARITY_SWEEPS, 10 ** 9), par_out)
    errs, known_ok, names = [], True, []
    for name, rl, rh, cl, ch in parity_crops(r0, r1, world, rank):
        err, ok = oracle_crop_check(par_out, r0, rl, rh, cl, ch)
        errs.append(err); known_ok = known_ok and ok; names.append(name)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (hashes, errs, known_ok, names))
        hashes = {k: v for p_ in parts for k, v in p_[0].items()}
        errs = [e for p_ in parts for e in p_[1]]
        known_ok = all(p_[2] for p_ in parts)
        names = [n for p_ in parts for n in p_[3]]
    want = None
    try:
        with open(HASH_FILE) as fh:
            want = json.load(fh)["blocks"]
    except Exception:
        pass
    if args.write_hashes and rank == 0 and world == 1 and S == SWEEPS_PER_STEP:
        with open(os.path.join(ROOT, "gpurun_out", os.path.basename(HASH_FILE)), "w") as fh:
            json.dump({"what": "SHA-256 (first 16 hex digits) of the 4096-row blocks of the CCST fp64 solve of the C5 grid "
                               "after %d sweeps on one GPU" % e2e_sweeps, "blocks": hashes}, fh, indent=1)
    parity = {"oracle_crops": names, "oracle_sweeps": PARITY_SWEEPS, "max_err_over_range": max(errs),
              "tolerance": 1e-10, "known_cells_bit_exact": known_ok,
              "block_sha256_equal_1gpu": (None if want is None or S != SWEEPS_PER_STEP else hashes == want),
              "pass": bool(max(errs) <= 1e-10 and known_ok and (want is None or S != SWEEPS_PER_STEP or hashes == want))}

    # ------------------------------------------------------------------ a second line of evidence: harmonic, same grid
    extra = None
    if world == 1 and not args.no_extra:
        p = make_params(GRID, GRID, dtype, C.HMI_HARMONIC, orientation=1, dt=0.2, device=local)
        with Solver(p) as hs:
            hs.set_problem_host(a_img, a_mask)
            hs.sweeps(120)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            hs.sweeps(600)
            ev1.record()
            torch.cuda.synchronize()
            extra = {"harmonic_f64_resident_glups": float(GRID) * GRID * 600 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9,
                     "harmonic_temporal_block": hs.temporal_block}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        g, k, secs, cores = cpu_port_glups()
        cpu = {"value": g, "unit": "GLUPS", "cores": cores, "kind": "port",
               "sample": "%d CCST sweeps (one call, %.1f s) of rows [0, %d) x %d columns of the same fp64 grid, OpenMP C "
                         "port of the reference loop (oracle/)" % (k, secs, CPU_BAND_ROWS, GRID)}

    if rank == 0:
        nlaunch = max(int(timing["launches"]), 1)
        traffic = ncu_traffic("ccst_f64_k%d_%d" % (tb, GRID))       # one launch over the whole grid (N = 1 capture)
        if traffic is not None and world > 1:
            traffic = int(traffic / world)                        # a slab launch moves its share of it
        achieved = float(GRID) * GRID * S * steps / world / nlaunch * BYTES_PER_LUP[dkey] / (ms * 1e-3 / nlaunch) / 1e9
        out = {
            "metric": "grid-cell updates/s", "value": value, "unit": "GLUPS", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dkey,
            "data": "synthetic",
            "config": {"workload": WORKLOAD + ("" if world == 1 else ", cut into %d row slabs of %d rows" % (world, GRID // world)),
                       "sweeps_per_step": S, "kernel": kname, "temporal_block": tb,
                       "exchange_every": None if world == 1 else E, "halo": None if world == 1 else args.halo,
                       "l2": "grid (2 x %d MiB + mask per GPU) larger than the 126 MB L2" % (GRID // world * GRID * 8 >> 20),
                       "extra": extra},
            "e2e": {"value": e2e_val, "unit": "GLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "sweeps_per_step": e2e_sweeps, "copies_overlapped": max(overlapped, 0) if world == 1 else 0,
                    "api": "CCSTInpainter(**opts).inpaint(image, mask, out=pinned)" if world == 1
                           else "distributed.inpaint_slab(CCSTInpainter(**opts), rows, mask, global_rows=..., out=pinned) per rank"},
            "gpu_launches": int(timing["launches"]),
            "clocks": timing["clocks"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "note": "per GPU: algorithmic 17 B per grid-cell update x %d sweeps fused per launch; temporal "
                                 "blocking lets achieved exceed the HBM peak" % tb},
            "parity": parity,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=SWEEPS_PER_STEP, help="sweeps per step")
    ap.add_argument("--temporal-block", type=int, default=0, help="sweeps fused per launch (0 = auto)")
    ap.add_argument("--exchange-every", type=int, default=12, help="sweeps between halo exchanges (N > 1)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="how ghost rows travel between GPUs")
    ap.add_argument("--vec", type=int, default=0, help="fp64 cells per lane of the streaming kernel (tuning)")
    ap.add_argument("--tma", type=int, default=0, help="row staging of the streaming kernel (tuning)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--write-hashes", action="store_true", help="write the block hashes of the 1-GPU run to gpurun_out/")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
