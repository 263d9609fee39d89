"""The reference's `convolver` option, reduced to what it means on the device.

In the reference a Convolver object is a strategy that runs one 3x3 filter on the host
(heightmap_interpolation/inpainting/convolver.py:8-69).  The five names are five different discrete
operators; at the cells a solve actually updates they differ only in orientation
(correlation vs true convolution) and border rule, so that is all this class carries.
"""
from .. import _capi as C

# name -> (orientation, border).  "scipy-ndimage" maps to the zero-padded SciPy-signal back end,
# exactly as the reference's factory does (convolver.py:17-18).
_TABLE = {
    "opencv": (+1, C.HMI_BORDER_REFLECT101),
    "masked": (-1, C.HMI_BORDER_REFLECT101),
    "masked-parallel": (-1, C.HMI_BORDER_REFLECT101),
    "scipy-signal": (-1, C.HMI_BORDER_ZERO),
    "scipy-ndimage": (-1, C.HMI_BORDER_ZERO),
}


class Convolver(object):
    def __init__(self, method):
        key = method.lower()
        if key not in _TABLE:
            raise ValueError("Unknown convolver type '{:s}'".format(method))
        self.method = key
        self.orientation, self.border = _TABLE[key]

    def __call__(self, image, kernel, mask):
        raise NotImplementedError("3x3 filters are fused into the CUDA sweep kernels of the B200 back-end")
