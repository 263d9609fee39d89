"""Harmonic / Laplacian inpainter (reference: inpainting/sobolev_inpainter.py:10-17, step = L f)."""
from .. import _capi as C
from .fd_pde_inpainter import FDPDEInpainter


class SobolevInpainter(FDPDEInpainter):
    METHOD = C.HMI_HARMONIC

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
