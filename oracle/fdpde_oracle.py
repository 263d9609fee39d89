"""TEST INFRASTRUCTURE ONLY — ctypes front end of the CPU oracle (oracle/fdpde_oracle.c).

Mirrors the reference's ``Inpainter(**options).inpaint(image, mask)`` surface
(/root/reference/heightmap_interpolation/inpainting/fd_pde_inpainter.py:52-185) closely enough
for parity tests to read like the reference's own usage.  Parity status: PINNED by
tests/test_oracle_golden.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).

Only tests/, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this
module.  The product package ``heightmap_interpolation_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfdpde_oracle.so")

METHODS = {"harmonic": 0, "sobolev": 0, "ccst": 1, "tv": 2, "amle": 3}
# convolver name -> (orientation, border, masked)       convolver.py:47-69, SURVEY.md A.2
CONVOLVERS = {
    "opencv": (+1, 0, 0),
    "masked": (-1, 0, 1),
    "masked-parallel": (-1, 0, 1),
    # convolver.py:17-18 maps "scipy-ndimage" to ScipySignalConvolver (zero padding); the
    # symmetric-border ScipyNDImageConvolver (:52-54) is unreachable by name.
    "scipy-ndimage": (-1, 2, 0),
    "scipy-signal": (-1, 2, 0),
}
CRITERIA = {"relative": 0, "absolute": 1, "absolute_percent": 2}


class _Params(ctypes.Structure):
    _fields_ = [
        ("method", ctypes.c_int),
        ("orientation", ctypes.c_int),
        ("border", ctypes.c_int),
        ("masked", ctypes.c_int),
        ("criteria", ctypes.c_int),
        ("reserved", ctypes.c_int),
        ("dt", ctypes.c_double),
        ("tension", ctypes.c_double),
        ("epsilon", ctypes.c_double),
        ("relaxation", ctypes.c_double),
        ("term_thres", ctypes.c_double),
        ("term_check_iters", ctypes.c_longlong),
        ("max_iters", ctypes.c_longlong),
    ]


def build(force=False):
    """Compile the oracle with gcc (building the checker is not using it)."""
    src = [os.path.join(_HERE, f) for f in ("fdpde_oracle.c", "fdpde_oracle_impl.h", "fdpde_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        for suf in ("f64", "f32"):
            g = getattr(L, "oracle_inpaint_grid_" + suf)
            g.restype = ctypes.c_int
            g.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                          ctypes.POINTER(_Params), ctypes.c_void_p,
                          ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_int)]
            s = getattr(L, "oracle_step_fun_" + suf)
            s.restype = ctypes.c_int
            s.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                          ctypes.POINTER(_Params), ctypes.c_void_p]
        L.oracle_dilate3x3.restype = None
        L.oracle_dilate3x3.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _suffix(a):
    if a.dtype == np.float64:
        return "f64"
    if a.dtype == np.float32:
        return "f32"
    raise TypeError("oracle supports float32/float64 grids, got %s" % a.dtype)


class OracleInpainter:
    """CPU oracle with the reference's constructor options."""

    def __init__(self, method, **kwargs):
        self.method = method.lower()
        if self.method not in METHODS:
            raise ValueError("unknown method " + method)
        self.dt = kwargs.pop("update_step_size", 0.01)
        self.term_check_iters = kwargs.pop("term_check_iters", 1000)
        self.term_thres = kwargs.pop("term_thres", 1e-8)
        self.max_iters = int(kwargs.pop("max_iters", 1e8))
        self.relaxation = kwargs.pop("relaxation", 0)
        self.init_with = kwargs.pop("init_with", "zeros")
        self.convolver_type = kwargs.pop("convolver", "masked")
        self.term_criteria = kwargs.pop("term_criteria", "relative")
        self.tension = kwargs.pop("tension", 0.0)
        self.epsilon = kwargs.pop("epsilon", 1e-2)
        # amle_inpainter.py:26-37: scipy.ndimage.convolve1d (mode='reflect') on the whole grid
        self.convolve_in_1d = kwargs.pop("convolve_in_1d", False)
        self.mgs_levels = kwargs.pop("mgs_levels", 1)
        self.mgs_min_res = kwargs.pop("mgs_min_res", 100)
        if self.convolver_type.lower() not in CONVOLVERS:
            raise ValueError("Unknown convolver type '{:s}'".format(self.convolver_type))
        # results of the last solve (the reference only prints these)
        self.updates_done = None
        self.last_diff = None
        self.converged = None

    def _params(self, term_thres, max_iters=None):
        o, b, m = CONVOLVERS[self.convolver_type.lower()]
        if self.convolve_in_1d and self.method == "amle":
            o, b, m = -1, 1, 0          # true convolution, symmetric border, whole grid
        crit = CRITERIA.get(self.term_criteria, -1)
        if crit < 0:
            raise ValueError("Unknown termination criteria!")
        return _Params(METHODS[self.method], o, b, m, 0 if crit == 0 else 1, 0,
                       float(self.dt), float(self.tension), float(self.epsilon),
                       float(self.relaxation), float(term_thres),
                       int(self.term_check_iters), int(self.max_iters if max_iters is None else max_iters))

    def initialize(self, image, mask):
        """initializer.py:14-18 — in place on the caller's array, zeros/mean only."""
        if self.init_with == "zeros":
            image[~mask] = 0
        elif self.init_with == "mean":
            image[~mask] = np.mean(image[mask])
        elif self.init_with == "nearest":
            nearest_init(image, mask)
        else:
            raise ValueError("oracle supports init_with zeros/mean/nearest only")
        return image

    def step_fun(self, f, work_mask=None):
        f = np.ascontiguousarray(f)
        suf = _suffix(f)
        out = np.empty_like(f)
        wm = None
        if work_mask is not None:
            wm = np.ascontiguousarray(work_mask, dtype=np.uint8)
        p = self._params(self.term_thres)
        rc = getattr(lib(), "oracle_step_fun_" + suf)(
            f.ctypes.data, None if wm is None else wm.ctypes.data, f.shape[0], f.shape[1],
            ctypes.byref(p), out.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_step_fun failed rc=%d" % rc)
        return out

    def inpaint_grid(self, image, mask):
        image = np.ascontiguousarray(image)
        suf = _suffix(image)
        m8 = np.ascontiguousarray(mask, dtype=np.uint8)
        thres = self.term_thres
        if self.term_criteria == "absolute_percent":      # fd_pde_inpainter.py:149-155 (mutates self)
            z_range = abs(np.max(image) - np.min(image))    # stays in the grid dtype, like NumPy does
            self.term_thres = z_range * self.term_thres
            thres = self.term_thres
        p = self._params(thres)
        out = np.empty_like(image)
        done = ctypes.c_longlong(0)
        diff = ctypes.c_double(0)
        conv = ctypes.c_int(0)
        rc = getattr(lib(), "oracle_inpaint_grid_" + suf)(
            image.ctypes.data, m8.ctypes.data, image.shape[0], image.shape[1], ctypes.byref(p),
            out.ctypes.data, ctypes.byref(done), ctypes.byref(diff), ctypes.byref(conv))
        if rc != 0:
            raise RuntimeError("oracle_inpaint_grid failed rc=%d" % rc)
        self.updates_done, self.last_diff, self.converged = done.value, diff.value, bool(conv.value)
        return out

    def inpaint(self, image, mask):
        if getattr(self, "mgs_levels", 1) > 1:
            return self.inpaint_multigrid(image, mask)
        image = self.initialize(image, mask)
        return self.inpaint_grid(image, mask)

    def inpaint_multigrid(self, image, mask):
        """fd_pde_inpainter.py:187-279 restated: stride-2 pyramid of views (misc/image_proc.py:4-12), the
        initialiser on the coarsest level only (through the view, i.e. into the caller's array),
        coarsest threshold orig * 2**num_levels, then per finer level cv2.resize (bilinear, restated in
        `resize_bilinear` below), the arithmetic blend up*(~mask) + image*mask and orig * 2**level.
        Records (rows, cols, updates) per level in self.history."""
        import math
        self.history = []
        if image.shape[0] < self.mgs_min_res or image.shape[1] < self.mgs_min_res:
            out = self.inpaint_grid(image, mask)            # :196-199: no initialiser on this path
            self.history.append((image.shape[0], image.shape[1], self.updates_done))
            return out
        images, masks = [image], [mask]
        num_levels = self.mgs_levels
        for level in range(1, self.mgs_levels):
            width = math.ceil(images[level - 1].shape[1] / 2)
            height = math.ceil(images[level - 1].shape[0] / 2)
            if width < self.mgs_min_res or height < self.mgs_min_res:
                num_levels = level
                break
            img_rs = images[level - 1][0::2, 0::2]
            images.append(img_rs)
            m_rs = np.asarray(masks[level - 1], dtype="uint8")[0::2, 0::2]
            masks.append(np.logical_and(m_rs, ~np.isnan(img_rs)))
        init = self.initialize(images[num_levels - 1], masks[num_levels - 1])
        original = self.term_thres
        self.term_thres = original * 2 ** num_levels
        coarse = self.inpaint_grid(init, masks[num_levels - 1] > 0)
        self.history.append((init.shape[0], init.shape[1], self.updates_done))
        if num_levels == 1:
            return coarse
        result = coarse
        for level in range(num_levels - 2, -1, -1):
            self.term_thres = original * 2 ** level
            img, m = images[level], masks[level]
            up = resize_bilinear(coarse, img.shape[0], img.shape[1])
            if level == 0 and np.any(np.isnan(img)):
                img[np.isnan(img)] = 0
            blended = up * (~m) + img * m
            result = self.inpaint_grid(blended, m)
            self.history.append((img.shape[0], img.shape[1], self.updates_done))
            if level > 0:
                coarse = result
        return result


def nearest_init(image, mask, return_d2=False):
    """initializer.py:19-34 with method='nearest' — scipy.interpolate.griddata over all known cells:
    every unknown cell takes the value of its Euclidean-nearest known cell, in place.

    Brute force in NumPy (small grids only).  SciPy's cKDTree returns *a* nearest point; which one
    among equidistant points depends on its leaf order, so the restatement fixes a rule (and the
    CUDA path follows the same one): smallest distance, then smallest |column offset|, then the
    left column, then the upper row.  Pinned against the reference by tests/test_oracle_golden.py:
    identical wherever the nearest cell is unique, same distance elsewhere."""
    mask = np.asarray(mask, dtype=bool)
    ky, kx = np.nonzero(mask)
    if ky.size == 0:
        raise ValueError("the nearest initialiser needs at least one known cell")
    vals = image[mask]
    d2 = np.zeros(image.shape, dtype=np.int64)
    uy, ux = np.nonzero(~mask)
    for i, j in zip(uy, ux):
        dj = kx - j
        dd = (ky - i) ** 2 + dj ** 2
        # lexicographic key: distance, |dj|, left before right, upper before lower
        key = np.lexsort((ky, dj, np.abs(dj), dd))[0]
        image[i, j] = vals[key]
        d2[i, j] = dd[key]
    return (image, d2) if return_d2 else image


def resize_bilinear(src, rows, cols):
    """cv2.resize(src, (cols, rows)) with the default INTER_LINEAR (fd_pde_inpainter.py:255 passes no
    interpolation argument): half-pixel centres, fx = (d + 0.5) * (ssize / dsize) - 0.5 clamped at both
    ends, weights rounded to the image's own precision, columns first and then rows.  (OpenCV is a
    third-party dependency absent from /root/reference; this is its published algorithm, and the
    multigrid fixtures produced by the reference pin it.)"""
    src = np.asarray(src)
    wt = np.float32 if src.dtype == np.float32 else np.float64

    def taps(ssize, dsize):
        d = np.arange(dsize, dtype=np.float64)
        f = (d + 0.5) * (float(ssize) / float(dsize)) - 0.5
        i0 = np.floor(f).astype(np.int64)
        f = f - i0
        f[i0 < 0] = 0.0
        i0[i0 < 0] = 0
        hi = i0 >= ssize - 1
        f[hi] = 0.0
        i0[hi] = ssize - 1
        i1 = np.minimum(i0 + 1, ssize - 1)
        f = f.astype(wt)
        return i0, i1, (wt(1.0) - f).astype(wt), f

    x0, x1, a0, a1 = taps(src.shape[1], cols)
    y0, y1, b0, b1 = taps(src.shape[0], rows)
    s = src.astype(wt, copy=False)
    tmp = s[:, x0] * a0[None, :] + s[:, x1] * a1[None, :]
    out = tmp[y0, :] * b0[:, None] + tmp[y1, :] * b1[:, None]
    return out.astype(src.dtype, copy=False)


def dilate3x3(unknown):
    u = np.ascontiguousarray(unknown, dtype=np.uint8)
    out = np.empty_like(u)
    lib().oracle_dilate3x3(u.ctypes.data, out.ctypes.data, u.shape[0], u.shape[1])
    return out
