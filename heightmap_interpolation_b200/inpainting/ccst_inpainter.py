"""Continuous Curvature Splines in Tension inpainter (reference: inpainting/ccst_inpainter.py:9-39).

step = -((1 - t) L(L f) - t L f); tension = 0 is the biharmonic inpainter, tension = 1 the harmonic one.
"""
from .. import _capi as C
from .fd_pde_inpainter import FDPDEInpainter


class CCSTInpainter(FDPDEInpainter):
    METHOD = C.HMI_CCST

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.tension = kwargs.pop("tension", 0.0)
        if self.tension < 0. or self.tension > 1.:
            raise ValueError("tension parameter must be a number between 0 and 1 (included)")

    def _method_params(self):
        return {"tension": self.tension}

    def get_config(self):
        config = super(CCSTInpainter, self).get_config()
        config["tension"] = self.tension
        return config
