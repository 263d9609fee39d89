"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Every rank solves its row slab with NCCL halo exchange; rank 0 gathers the slabs and compares with
the single-GPU solve of the whole grid (bit-exact expected: same kernels, same per-cell arithmetic)
and with the CPU oracle (tolerance 1e-10)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heightmap_interpolation_b200 as hmi                               # noqa: E402
from heightmap_interpolation_b200 import synthetic                       # noqa: E402
from heightmap_interpolation_b200.distributed import inpaint_slab        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [("harmonic", hmi.SobolevInpainter, dict(update_step_size=0.2), np.float64, 8),
             ("ccst", hmi.CCSTInpainter, dict(update_step_size=0.01, tension=0.3), np.float64, 6),
             ("ccst", hmi.CCSTInpainter, dict(update_step_size=0.01, tension=0.3), np.float32, 8),
             ("tv", hmi.TVInpainter, dict(update_step_size=0.225, epsilon=1.0), np.float64, 4)]
    rows, cols = 1024 + 37, 777
    for name, cls, opts, dtype, E in cases:
        img = synthetic.bathymetry((rows, cols), dtype=dtype)
        mask = synthetic.mask_random((rows, cols), 0.4, seed=3)
        img = synthetic.zero_unknown(img, mask)
        common = dict(convolver="opencv", term_criteria="relative", term_thres=1e-4, term_check_iters=25,
                      max_iters=200, device=local, **opts)
        inp = cls(**common)
        out, (r0, r1) = inpaint_slab(inp, img, mask, exchange_every=E)
        parts = [None] * world
        dist.all_gather_object(parts, (r0, r1, out, inp.iterations_, inp.last_diff_, inp.converged_))
        if rank == 0:
            got = np.vstack([p[2] for p in sorted(parts, key=lambda t: t[0])])
            single = cls(**common)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                want = single.inpaint(img.copy(), mask)
            same_iters = all(p[3] == single.iterations_ for p in parts)
            exact = np.array_equal(got, want)
            err = float(np.abs(got.astype(np.float64) - want).max()) / float(want.max() - want.min())
            print("%-8s %s world=%d E=%d iterations %s (single %d) converged %s bit-exact=%s err=%.2e exchanges=%d"
                  % (name, np.dtype(dtype).name, world, E, [p[3] for p in parts], single.iterations_,
                     parts[0][5], exact, err, inp.exchanges_), flush=True)
            ok = ok and same_iters and err <= (1e-10 if dtype == np.float64 else 1e-5)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
