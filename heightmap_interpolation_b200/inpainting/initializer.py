"""Initialisers for the unknown cells (reference: heightmap_interpolation/inpainting/initializer.py:6-44).

`zeros` and `mean` are what the hot path needs and run on the host in place, exactly as the
reference does (the arrays are still on the host at this point).  `nearest`/`linear`/`cubic` are
scattered-data interpolation over *all* known cells (scipy.interpolate.griddata, :19-34) and the
`sobolev` initialiser is a nested harmonic solve (:35-41); the latter is routed to the B200 harmonic
inpainter, the former stay SciPy calls and are outside the accelerated path.
"""
import numpy as np


class Initializer():

    def __init__(self, init_with):
        self.init_with = init_with
        available_options = ["zeros", "mean", "nearest", "linear", "cubic", "sobolev"]
        if self.init_with.lower() not in available_options:
            raise ValueError("init_with must be one of these options: " + ", ".join(available_options))

    def initialize(self, image, mask):
        kind = self.init_with.lower()
        if kind == "zeros":
            _fill_unknown(image, mask, 0.0)
        elif kind == "mean":
            _fill_unknown(image, mask, None)
        elif kind in ("nearest", "linear", "cubic"):
            from scipy import interpolate
            fill_value = 0 if kind == "nearest" else np.mean(image[mask])
            x = np.arange(0, image.shape[1])
            y = np.arange(0, image.shape[0])
            xx, yy = np.meshgrid(x, y)
            interp = interpolate.griddata((xx[mask], yy[mask]), image[mask].ravel(), (xx, yy),
                                          method=kind, fill_value=fill_value)
            image[~mask] = interp[~mask]
        else:  # "sobolev": harmonic solve with the reference's hard-coded options (:36-41)
            from .sobolev_inpainter import SobolevInpainter
            inpainter = SobolevInpainter(update_step_size=0.8 / 4, max_iters=1e5)
            image = inpainter.inpaint(image, mask)
            print("initialized with sobolev")
        return image


def _fill_unknown(image, mask, value):
    """image[~mask] = value (value None = mean of the known cells), in place.  Contiguous float
    grids go through the library's multi-threaded host helper; anything else uses NumPy."""
    import ctypes

    from .. import _capi as C
    fast = (isinstance(image, np.ndarray) and image.flags.c_contiguous and image.flags.writeable
            and image.dtype in (np.float32, np.float64) and isinstance(mask, np.ndarray)
            and mask.dtype == np.bool_ and mask.flags.c_contiguous and mask.shape == image.shape)
    if not fast:
        image[~mask] = np.mean(image[mask]) if value is None else value
        return
    code = C.HMI_F64 if image.dtype == np.float64 else C.HMI_F32
    lib = C.lib()
    if value is None:
        m = ctypes.c_double()
        C.check(lib.hmi_host_known_mean(image.ctypes.data, mask.ctypes.data, image.size, code,
                                        ctypes.byref(m)))
        value = image.dtype.type(np.mean(image[mask])) if image.size < (1 << 16) else m.value
    C.check(lib.hmi_host_fill_unknown(image.ctypes.data, mask.ctypes.data, image.size, code, float(value)))
