"""Absolutely Minimizing Lipschitz Extension inpainter (reference: inpainting/amle_inpainter.py:9-59).

step = uxx v0^2 + uyy v1^2 + (uxy + uyx) v0 v1 with v = Du/|Du| (2-D convolver branch, :41-59).
`convolve_in_1d=True` selects the reference's scipy.ndimage.convolve1d branch (:26-37), i.e. the
convolution orientation with symmetric borders on the whole grid.
"""
from .. import _capi as C
from .fd_pde_inpainter import FDPDEInpainter


class AMLEInpainter(FDPDEInpainter):
    METHOD = C.HMI_AMLE

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.convolve_in_1d = kwargs.pop("convolve_in_1d", False)

    def _method_params(self):
        if self.convolve_in_1d:
            return {"orientation": -1, "border": C.HMI_BORDER_SYMMETRIC}
        return {}
