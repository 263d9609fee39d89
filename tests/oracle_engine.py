"""TEST INFRASTRUCTURE: a CPU stand-in for CudaSlabEngine built on the oracle, so the host-side
slab logic (partitioning, halo exchange, trapezoid ghost advance, norm all-reduce) can be exercised
with gloo on CPU.  Never imported by the product package."""
import numpy as np
import torch

from heightmap_interpolation_b200 import _capi as C
from heightmap_interpolation_b200.solver import Norm
from oracle.fdpde_oracle import OracleInpainter

_METHOD = {C.HMI_HARMONIC: "harmonic", C.HMI_CCST: "ccst", C.HMI_TV: "tv", C.HMI_AMLE: "amle"}


class OracleSlabEngine:
    """Runs the oracle on the slab array *including* its ghost rows as if it were a whole grid.
    The wrong border rule at a seam only contaminates radius*n ghost rows per n sweeps — exactly the
    rows the trapezoid scheme gives up — so the interior rows stay exact as long as the caller
    refreshes the halos every ghost/radius sweeps."""

    def __init__(self, params):
        self.p = params
        self.device = torch.device("cpu")
        self.kernel_name = "oracle"
        conv = "opencv" if params.orientation == 1 else "masked"
        if params.border == C.HMI_BORDER_ZERO:
            conv = "scipy-signal"
        self.opts = dict(update_step_size=params.dt, tension=params.tension, epsilon=params.epsilon,
                         convolver=conv)
        self.method = _METHOD[params.method]

    def load(self, image, mask):
        self.f = torch.from_numpy(np.array(image, copy=True))
        self.mask = np.array(mask, copy=True)

    def rows_view(self, lo, hi):
        return self.f[lo:hi]

    def _run(self, n):
        o = OracleInpainter(self.method, max_iters=n, term_thres=1e-300, term_check_iters=10 ** 9,
                            **self.opts)
        return o.inpaint_grid(self.f.numpy(), self.mask)

    def sweeps(self, n):
        if n > 0:
            self.f.copy_(torch.from_numpy(self._run(n)))

    def sweeps_norm(self, n):
        if n > 1:
            self.sweeps(n - 1)
        gt, R = self.p.ghost_top, self.p.rows
        prev = self.f.numpy()[gt:gt + R].copy()
        self.sweeps(1)
        new = self.f.numpy()[gt:gt + R]
        d = new - prev
        return Norm(float(np.sum(d * d)), float(np.sum(new * new)), float(np.abs(d).max()), 0.0)

    def result(self):
        gt, R = self.p.ghost_top, self.p.rows
        return self.f.numpy()[gt:gt + R].copy()

    def close(self):
        pass
