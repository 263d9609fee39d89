from .sobolev_inpainter import SobolevInpainter
from .ccst_inpainter import CCSTInpainter
from .tv_inpainter import TVInpainter
from .amle_inpainter import AMLEInpainter
from .fd_pde_inpainter import FDPDEInpainter

__all__ = ["FDPDEInpainter", "SobolevInpainter", "CCSTInpainter", "TVInpainter", "AMLEInpainter"]
