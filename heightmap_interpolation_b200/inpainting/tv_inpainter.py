"""Total-variation inpainter (reference: inpainting/tv_inpainter.py:7-38).

step = div(grad f / sqrt(|grad f|^2 + eps^2)) with forward differences for the gradient and
backward differences for the divergence; the result depends on the convolver's orientation.
"""
from .. import _capi as C
from .fd_pde_inpainter import FDPDEInpainter


class TVInpainter(FDPDEInpainter):
    METHOD = C.HMI_TV

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = kwargs.pop("epsilon", 1e-2)
        if self.epsilon <= 0:
            raise ValueError("Epsilon parameter must be greater than zero")

    def _method_params(self):
        return {"epsilon": self.epsilon}

    def get_config(self):
        config = super(TVInpainter, self).get_config()
        config["epsilon"] = self.epsilon
        return config
