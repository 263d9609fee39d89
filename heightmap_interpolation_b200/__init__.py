"""B200-native FD-PDE inpainting: a drop-in for the explicit finite-difference PDE solvers of
coronis-computing/heightmap_interpolation (heightmap_interpolation/inpainting/), running as
hand-written CUDA kernels for sm_100a behind the C ABI in include/hmi_b200.h."""
from .inpainting import (AMLEInpainter, CCSTInpainter, FDPDEInpainter, SobolevInpainter,
                         TVInpainter)

__all__ = ["FDPDEInpainter", "SobolevInpainter", "CCSTInpainter", "TVInpainter", "AMLEInpainter"]
__version__ = "0.1.0"
