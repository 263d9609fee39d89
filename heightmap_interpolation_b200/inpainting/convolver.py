"""The reference's `convolver` option, reduced to what it means on the device.

In the reference a Convolver object is a strategy that runs one 3x3 filter on the host
(heightmap_interpolation/inpainting/convolver.py:8-69).  The five names are five different discrete
operators; at the cells a solve actually updates they differ only in orientation
(correlation vs true convolution) and border rule, so that is all the solver needs of this class.
Calling the object filters one array on the GPU (hmi_convolve3x3_host), with the reference's
signature, for code written against that hook.
"""
import ctypes

import numpy as np

from .. import _capi as C

# the convolvers that evaluate only where `mask` is set and copy their input elsewhere
# (convolver.py:62-69 -> convolve_at_mask.py:39-63); the others ignore the mask argument (:47-59)
_MASKED = ("masked", "masked-parallel")

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

        self.device = 0

    def __call__(self, image, kernel, mask):
        """One 3x3 filter of `image` (2-D float32 / float64) on the GPU, like the reference's call.  For
        float32 grids each output is accumulated in double and rounded once (what the numba back ends do;
        OpenCV accumulates in float32, so "opencv" may differ from the reference in the last bit)."""
        image = np.ascontiguousarray(image)
        if image.ndim != 2 or image.dtype not in (np.float32, np.float64):
            raise TypeError("image must be a 2-D float32 or float64 array")
        k = np.ascontiguousarray(kernel, dtype=np.float64)
        if k.shape != (3, 3):
            raise ValueError("the B200 convolver filters with 3x3 kernels (the reference's stencils)")
        m8 = None
        if self.method in _MASKED:
            if mask is None:
                raise ValueError("the masked convolvers need a mask")
            m8 = np.ascontiguousarray(mask, dtype=np.uint8)
            if m8.shape != image.shape:
                raise ValueError("mask must have the same shape as the image")
        out = np.empty_like(image)
        C.check(C.lib().hmi_convolve3x3_host(
            image.shape[0], image.shape[1], C.HMI_F64 if image.dtype == np.float64 else C.HMI_F32, int(self.device),
            image.ctypes.data, image.strides[0], k.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), self.orientation,
            self.border, None if m8 is None else m8.ctypes.data, 0 if m8 is None else m8.strides[0],
            out.ctypes.data, out.strides[0]))
        return out
