"""Initialisers for the unknown cells (reference: heightmap_interpolation/inpainting/initializer.py:6-44).

`zeros` and `mean` run on the host in place, exactly as the reference does (the arrays are still on
the host at this point).  `nearest` — the CLI default (apps/apps_common.py:52) — is a cKDTree query
over *all* known cells in the reference (scipy.interpolate.griddata, :19-34), 3 s of host time on a
2048^2 grid; here it is an exact Euclidean nearest-known-cell search on the B200
(csrc/hmi_init.cu, `hmi_init_nearest_host`), identical to SciPy wherever the nearest cell is unique
and equidistant otherwise.  `linear`/`cubic` need a Delaunay triangulation of all known cells and
stay SciPy calls, outside the accelerated path.  The `sobolev` initialiser is a
nested harmonic solve (:35-41) and is routed to the B200 harmonic inpainter.

Extra keyword ``nearest_backend`` ("b200" | "scipy") selects the SciPy implementation of `nearest`
for comparisons.
"""
import numpy as np


class Initializer():

    def __init__(self, init_with, nearest_backend="b200", device=0):
        self.init_with = init_with
        self.nearest_backend = nearest_backend
        self.device = device
        available_options = ["zeros", "mean", "nearest", "linear", "cubic", "sobolev"]
        if self.init_with.lower() not in available_options:
            raise ValueError("init_with must be one of these options: " + ", ".join(available_options))

    def initialize(self, image, mask):
        kind = self.init_with.lower()
        if kind == "zeros":
            _fill_unknown(image, mask, 0.0)
        elif kind == "mean":
            _fill_unknown(image, mask, None)
        elif kind == "nearest" and self.nearest_backend == "b200":
            _nearest_b200(image, mask, self.device)
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


def _nearest_b200(image, mask, device=0):
    """image[~mask] = value of the Euclidean-nearest known cell, in place (hmi_init_nearest_host).
    Non-contiguous arrays (the strided pyramid views of the multigrid driver) go through a
    contiguous copy and are written back, so the caller's array is updated like the reference's."""
    import ctypes

    from .. import _capi as C
    if not isinstance(image, np.ndarray) or image.ndim != 2:
        raise ValueError("image must be a 2-D array")
    mask = np.asarray(mask)
    if mask.shape != image.shape:
        raise ValueError("mask must have the same shape as the image")
    if image.dtype not in (np.float32, np.float64):
        raise TypeError("image dtype must be float32 or float64, got %s" % image.dtype)
    m8 = np.ascontiguousarray(mask != 0 if mask.dtype != np.bool_ else mask).view(np.uint8)
    direct = image.flags.c_contiguous and image.flags.writeable
    work = image if direct else np.ascontiguousarray(image)
    code = C.HMI_F64 if image.dtype == np.float64 else C.HMI_F32
    C.check(C.lib().hmi_init_nearest_host(work.ctypes.data, work.strides[0], m8.ctypes.data, m8.strides[0],
                                          image.shape[0], image.shape[1], code, int(device or 0)))
    if not direct:
        unknown = m8.view(np.bool_) == False  # noqa: E712
        image[unknown] = work[unknown]
