#!/usr/bin/env python
"""Benchmark of the FD-PDE inpainting hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): grid-cell updates per second, GLUPS = rows*cols*sweeps / s / 1e9.

Workload at N = 1 (BASELINE.json configs[1], SURVEY.md 8d "C2"): CCST inpainting, tension 0.3,
dt 0.01, 4096x4096 synthetic bathymetry with survey-track gaps, fp64.  One *step* is
SWEEPS_PER_STEP relaxation sweeps of the whole grid.  At N > 1 (torchrun, one rank per GPU) the
grid grows to (4096*N) x 4096 rows, cut into N row slabs of the N = 1 size (weak scaling) with NCCL
halo exchange every `exchange_every` sweeps.

  value : whole-job GLUPS with the grid resident in HBM, timed with CUDA events on the launch stream
  e2e   : the same metric through the public API, Inpainter(**opts).inpaint(image, mask), with host
          buffers: pinned H2D of image+mask and D2H of the result inside the timed region
  roofline.achieved : algorithmic bytes (17 B fp64 / 9 B fp32 per grid-cell update)
          per launch / average launch duration; with temporal blocking (k sweeps per launch) this can
          exceed the HBM peak, so roofline.traffic carries the ncu-measured DRAM bytes per launch
  cpu_baseline : the CPU oracle (oracle/, an OpenMP C port of the reference loop) on the host cores

`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the
reference is pure Python + NumPy/OpenCV/Numba and cannot travel to the GPU box) on the same config.
"""
import argparse
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

GRID = 4096                # C2: 4096 x 4096
SWEEPS_PER_STEP = 1000     # one reference check window (term_check_iters default)
BYTES_PER_LUP = {"f64": 17, "f32": 9}
CCST = dict(update_step_size=0.01, tension=0.3, convolver="opencv")


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


def make_problem(rows, cols, dtype):
    from heightmap_interpolation_b200 import synthetic
    img = synthetic.bathymetry((rows, cols), seed=7, dtype=dtype)
    mask = synthetic.mask_tracks((rows, cols), 8, 16, 256)
    return synthetic.zero_unknown(img, mask), mask


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


def cpu_port_glups(img, mask, budget_s=15.0):
    """Time the CPU oracle on a bounded sample of the same workload: enough CCST sweeps of the same
    grid to fill ~budget_s on the host cores (all of them, OpenMP)."""
    from oracle.fdpde_oracle import OracleInpainter

    def run(k):
        o = OracleInpainter("ccst", max_iters=k, term_thres=1e-300, term_check_iters=10 ** 9, **CCST)
        t = time.perf_counter()
        o.inpaint_grid(img, mask)
        return time.perf_counter() - t
    run(1)                                  # warm-up (page faults, thread pool)
    t1 = run(2) / 2.0
    k = int(max(3, min(5000, budget_s / max(t1, 1e-6))))
    t = run(k)
    return img.size * k / t / 1e9, k, t


def run_reference(args):
    """Reference arm: the reference algorithm on the host CPU (oracle port), same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    img, mask = make_problem(GRID, GRID, np.float64)
    from oracle.fdpde_oracle import OracleInpainter
    sweeps = 20                              # bounded sample per step (~0.1-0.5 s on the box)
    o = OracleInpainter("ccst", max_iters=sweeps, term_thres=1e-300, term_check_iters=10 ** 9, **CCST)
    for _ in range(args.warmup):
        o.inpaint_grid(img, mask)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.inpaint_grid(img, mask)
    dt = time.perf_counter() - t0
    glups = img.size * sweeps * args.steps / dt / 1e9
    cores = os.cpu_count()
    sample = "%d CCST sweeps of the 4096x4096 fp64 grid per step (OpenMP C port of the reference loop)" % sweeps
    print(json.dumps({
        "impl": "reference", "metric": "grid-cell updates/s", "value": glups, "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CCST tension=0.3 dt=0.01, 4096x4096 synthetic bathymetry, "
                               "survey-track gaps (tracks 8/16/256), fp64",
                   "sweeps_per_step": sweeps},
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist

    import heightmap_interpolation_b200 as hmi
    from heightmap_interpolation_b200 import _capi as C
    from heightmap_interpolation_b200.solver import Solver, make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d"
                             % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dtype, dkey = np.float64, "f64"
    S = args.sweeps
    rows_global = GRID * world
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident throughput
    if world == 1:
        img, mask = make_problem(GRID, GRID, dtype)
        p = make_params(GRID, GRID, dtype, C.HMI_CCST, orientation=1, dt=CCST["update_step_size"],
                        tension=CCST["tension"], device=local, temporal_block=args.temporal_block)
        solver = Solver(p)
        solver.set_problem_host(img, mask)
        step = lambda: solver.sweeps(S)                                   # noqa: E731
        launches = lambda: solver.launch_count                            # noqa: E731
        kname, tb, exch = solver.kernel_name, solver.temporal_block, None
    else:
        from heightmap_interpolation_b200.distributed import CudaSlabEngine, SlabSolver, slab_bounds
        img_g, mask_g = make_problem(rows_global, GRID, dtype)
        r0, r1 = slab_bounds(rows_global, world)[rank]
        E = args.exchange_every
        ghost = 2 * E
        gt, gb = (ghost if rank > 0 else 0), (ghost if rank < world - 1 else 0)
        p = make_params(r1 - r0, GRID, dtype, C.HMI_CCST, orientation=1, dt=CCST["update_step_size"],
                        tension=CCST["tension"], device=local, temporal_block=args.temporal_block,
                        ghost_top=gt, ghost_bottom=gb)
        eng = CudaSlabEngine(p)
        eng.load(np.ascontiguousarray(img_g[r0 - gt:r1 + gb]), np.ascontiguousarray(mask_g[r0 - gt:r1 + gb]))
        slab = SlabSolver(eng, rank, world, 2, gt, gb, E, r1 - r0)
        img, mask = np.ascontiguousarray(img_g[r0:r1]), np.ascontiguousarray(mask_g[r0:r1])
        del img_g, mask_g
        step = lambda: slab.sweeps(S)                                     # noqa: E731
        launches = lambda: eng.solver.launch_count                        # noqa: E731
        kname, tb, exch = eng.solver.kernel_name, eng.solver.temporal_block, E

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    nlaunch = launches() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    lups = float(rows_global) * GRID * S * args.steps
    value = lups / (ms * 1e-3) / 1e9

    # ------------------------------------------------------------------ end to end through the public API
    e2e_sweeps = S + 1          # one default check window: checks at sweeps 1 and S+1
    if world == 1:
        pin_img = torch.from_numpy(img).pin_memory()
        pin_mask = torch.from_numpy(mask).pin_memory()
        pin_out = torch.empty_like(pin_img).pin_memory()
        a_img, a_mask, a_out = pin_img.numpy(), pin_mask.numpy(), pin_out.numpy()

        def e2e_step():
            inp = hmi.CCSTInpainter(term_check_iters=S, max_iters=e2e_sweeps, term_thres=1e-300,
                                    init_with="zeros", device=local,
                                    temporal_block=args.temporal_block, **CCST)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):      # "[WARNING] did NOT converge" by design
                inp.inpaint(a_img, a_mask, out=a_out)
            return inp.iterations_
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            done = e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        assert done == e2e_sweeps
        e2e_val = float(GRID) * GRID * e2e_sweeps / e2e_s / 1e9
        h2d = img.nbytes + mask.nbytes
        d2h = img.nbytes
    else:
        # slabs: per step every rank uploads its slab from pinned memory, runs, downloads its rows
        full_i, full_m = make_problem(rows_global, GRID, dtype)
        pin_img = torch.from_numpy(np.ascontiguousarray(full_i[r0 - gt:r1 + gb])).pin_memory()
        pin_mask = torch.from_numpy(np.ascontiguousarray(full_m[r0 - gt:r1 + gb])).pin_memory()
        del full_i, full_m
        pin_out = torch.empty((r1 - r0, GRID), dtype=torch.float64).pin_memory()

        def e2e_step():
            eng.load(pin_img.numpy(), pin_mask.numpy())
            slab.sweeps(e2e_sweeps)
            eng.solver.get_result_host(pin_out.numpy())
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item()) / args.steps
        e2e_val = float(rows_global) * GRID * e2e_sweeps / e2e_s / 1e9
        h2d = (pin_img.numel() * 8 + pin_mask.numel()) * world
        d2h = pin_out.numel() * 8 * world

    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        g, k, secs = cpu_port_glups(img, mask)
        cpu = {"value": g, "unit": "GLUPS", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d CCST sweeps of the same 4096x4096 fp64 grid in %.1f s (OpenMP C port of "
                         "the reference loop, oracle/)" % (k, secs)}

    if rank == 0:
        per_launch_lups = float(rows_global) * GRID * S * args.steps / max(nlaunch, 1) / world
        per_launch_s = ms * 1e-3 / max(nlaunch, 1)
        achieved = per_launch_lups * BYTES_PER_LUP[dkey] / per_launch_s / 1e9
        out = {
            "metric": "grid-cell updates/s", "value": value, "unit": "GLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dkey,
            "data": "synthetic",
            "config": {"workload": "CCST tension=0.3 dt=0.01, %dx%d synthetic bathymetry, survey-track "
                                   "gaps (tracks 8/16/256), fp64%s" % (
                                       rows_global, GRID,
                                       "" if world == 1 else ", %d row slabs of 4096 rows" % world),
                       "sweeps_per_step": S, "kernel": kname, "temporal_block": tb,
                       "exchange_every": exch,
                       "l2": "grid (2 x %d MiB + mask) larger than the 126 MB L2" % (GRID * GRID * 8 >> 20)},
            "e2e": {"value": e2e_val, "unit": "GLUPS", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "sweeps_per_step": e2e_sweeps,
                    "api": "CCSTInpainter(**opts).inpaint(image, mask, out=pinned)" if world == 1
                           else "slab load + sweeps + get_result_host per rank"},
            "gpu_launches": int(nlaunch),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic("ccst_f64_k%d" % tb),
                         "peak_source": peak_src,
                         "note": "algorithmic 17 B per grid-cell update x %d sweeps fused per launch; "
                                 "temporal blocking lets achieved exceed the HBM peak" % tb},
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=SWEEPS_PER_STEP, help="sweeps per step")
    ap.add_argument("--temporal-block", type=int, default=0, help="sweeps fused per launch (0 = auto)")
    ap.add_argument("--exchange-every", type=int, default=12, help="sweeps between halo exchanges (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
