"""Thin RAII wrapper over the C-ABI solver handle (include/hmi_b200.h)."""
import ctypes

import numpy as np

from . import _capi as C

_DTYPES = {np.dtype(np.float32): C.HMI_F32, np.dtype(np.float64): C.HMI_F64}


def dtype_code(dtype):
    try:
        return _DTYPES[np.dtype(dtype)]
    except KeyError:
        raise TypeError("grid dtype must be float32 or float64, got %s" % (dtype,))


def make_params(rows, cols, dtype, method, orientation=1, border=C.HMI_BORDER_REFLECT101,
                criteria=C.HMI_RELATIVE, dt=0.01, tension=0.0, epsilon=1e-2, relaxation=0.0,
                term_thres=1e-8, term_check_iters=1000, max_iters=int(1e8), device=0,
                kernel=C.HMI_KERNEL_AUTO, temporal_block=0, ghost_top=0, ghost_bottom=0):
    p = C.HmiParams()
    p.struct_size = ctypes.sizeof(C.HmiParams)
    p.device = int(device)
    p.rows, p.cols = int(rows), int(cols)
    p.dtype = dtype_code(dtype)
    p.method = int(method)
    p.orientation = int(orientation)
    p.border = int(border)
    p.criteria = int(criteria)
    p.kernel = int(kernel)
    p.temporal_block = int(temporal_block)
    p.ghost_top, p.ghost_bottom = int(ghost_top), int(ghost_bottom)
    p.dt, p.tension, p.epsilon = float(dt), float(tension), float(epsilon)
    p.relaxation = float(relaxation)
    p.term_thres = float(term_thres)
    p.term_check_iters = int(term_check_iters)
    p.max_iters = int(max_iters)
    return p


class Norm:
    """Convergence sums of one slab (hmi_norm)."""
    __slots__ = ("sum_d2", "sum_f2", "max_abs", "has_nan")

    def __init__(self, sum_d2=0.0, sum_f2=0.0, max_abs=0.0, has_nan=0.0):
        self.sum_d2, self.sum_f2, self.max_abs, self.has_nan = sum_d2, sum_f2, max_abs, has_nan

    def diff(self, criteria):
        """term_check(): fd_pde_inpainter.py:368-379."""
        if criteria == C.HMI_RELATIVE:
            with np.errstate(divide="ignore", invalid="ignore"):
                return float(np.sqrt(np.float64(self.sum_d2)) / np.sqrt(np.float64(self.sum_f2)))
        return float("nan") if self.has_nan else float(self.max_abs)


class Solver:
    """One grid (or one row slab) resident on one B200."""

    def __init__(self, params):
        self._lib = C.lib()
        self.params = params
        h = ctypes.c_void_p()
        C.check(self._lib.hmi_create(ctypes.byref(params), ctypes.byref(h)))
        self._h = h
        self.np_dtype = np.float64 if params.dtype == C.HMI_F64 else np.float32

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.hmi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- problem
    @property
    def phys_rows(self):
        p = self.params
        return p.ghost_top + p.rows + p.ghost_bottom

    def set_problem_host(self, image, mask):
        image = np.ascontiguousarray(image, dtype=self.np_dtype)
        m8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == np.bool_ else \
            np.ascontiguousarray(mask, dtype=np.uint8)
        want = (self.phys_rows, self.params.cols)
        if image.shape != want or m8.shape != want:
            raise ValueError("image/mask shape %s/%s, solver expects %s" % (image.shape, m8.shape, want))
        C.check(self._lib.hmi_set_problem_host(self._h, image.ctypes.data, image.strides[0],
                                               m8.ctypes.data, m8.strides[0]))

    def set_problem_host_blend(self, image, mask, coarse, xtab, ytab, scrub_nan=False):
        """Multigrid: load this (finer) level and blend the up-sampled current iterate of `coarse`
        into its unknown cells on the device (hmi_set_problem_host_blend).  xtab / ytab are the
        (src0, src1, w0, w1) tables of cv2's INTER_LINEAR per axis.  Returns HmiBlendInfo."""
        image = np.ascontiguousarray(image, dtype=self.np_dtype)
        m8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == np.bool_ else \
            np.ascontiguousarray(mask, dtype=np.uint8)
        want = (self.params.rows, self.params.cols)
        if image.shape != want or m8.shape != want:
            raise ValueError("image/mask shape %s/%s, solver expects %s" % (image.shape, m8.shape, want))
        tabs = []
        for (s0, s1, w0, w1), n in ((xtab, want[1]), (ytab, want[0])):
            tabs += [np.ascontiguousarray(s0, dtype=np.int32), np.ascontiguousarray(s1, dtype=np.int32),
                     np.ascontiguousarray(w0, dtype=self.np_dtype), np.ascontiguousarray(w1, dtype=self.np_dtype)]
            if any(t.shape != (n,) for t in tabs[-4:]):
                raise ValueError("interpolation tables must have one entry per output column / row")
        info = C.HmiBlendInfo()
        C.check(self._lib.hmi_set_problem_host_blend(
            self._h, image.ctypes.data, image.strides[0], m8.ctypes.data, m8.strides[0], coarse._h,
            *[t.ctypes.data for t in tabs], 1 if scrub_nan else 0, ctypes.byref(info)))
        return info

    def set_term_thres(self, term_thres):
        C.check(self._lib.hmi_set_term_thres(self._h, float(term_thres)))
        self.params.term_thres = float(term_thres)

    def set_problem_device(self, image_ptr, image_pitch, mask_ptr, mask_pitch, stream=0):
        C.check(self._lib.hmi_set_problem_device(self._h, image_ptr, image_pitch, mask_ptr,
                                                 mask_pitch, stream))

    # -- compute
    def sweeps(self, n, stream=0):
        C.check(self._lib.hmi_sweeps(self._h, int(n), stream))

    def sweeps_overlap(self, n, main_stream, side_stream):
        C.check(self._lib.hmi_sweeps_overlap(self._h, int(n), main_stream, side_stream))

    def sweeps_norm(self, n, stream=0):
        nm = C.HmiNorm()
        C.check(self._lib.hmi_sweeps_norm(self._h, int(n), stream, ctypes.byref(nm)))
        return Norm(nm.sum_d2, nm.sum_f2, nm.max_abs, nm.has_nan)

    # -- peer-memory halo links (row slabs)
    def halo_setup(self):
        """Allocate this slab's halo block; returns HmiHaloInfo (device pointer + 64-byte CUDA IPC handle)."""
        info = C.HmiHaloInfo()
        C.check(self._lib.hmi_halo_setup(self._h, ctypes.byref(info)))
        return info

    def halo_connect(self, side, peer_block, peer_device):
        C.check(self._lib.hmi_halo_connect(self._h, int(side), peer_block, int(peer_device)))

    def halo_connect_ipc(self, side, handle):
        handle = bytes(handle)
        if len(handle) != 64:
            raise ValueError("a CUDA IPC memory handle is 64 bytes")
        C.check(self._lib.hmi_halo_connect_ipc(self._h, int(side), handle))

    def sweeps_exchange(self, n, main_stream=0, side_stream=0):
        C.check(self._lib.hmi_sweeps_exchange(self._h, int(n), main_stream, side_stream))

    def run(self, stream=0):
        st = C.HmiStatus()
        C.check(self._lib.hmi_run(self._h, stream, ctypes.byref(st)))
        return st

    # -- results
    def get_result_host(self, out=None):
        p = self.params
        if out is None:
            out = np.empty((p.rows, p.cols), dtype=self.np_dtype)
            if out.nbytes >= (1 << 22):
                self._lib.hmi_host_prefault(out.ctypes.data, out.nbytes)
        C.check(self._lib.hmi_get_result_host(self._h, out.ctypes.data, out.strides[0]))
        return out

    def get_result_device(self, out_ptr, out_pitch, stream=0):
        C.check(self._lib.hmi_get_result_device(self._h, out_ptr, out_pitch, stream))

    def current_grid(self):
        ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        C.check(self._lib.hmi_current_grid(self._h, ctypes.byref(ptr), ctypes.byref(pitch)))
        return ptr.value, pitch.value

    def mask_grid(self):
        ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        C.check(self._lib.hmi_mask_grid(self._h, ctypes.byref(ptr), ctypes.byref(pitch)))
        return ptr.value, pitch.value

    # -- tuning / introspection
    def set_option(self, key, value):
        C.check(self._lib.hmi_set_option(self._h, key.encode(), int(value)))

    @property
    def temporal_block(self):
        return self._lib.hmi_temporal_block(self._h)

    @property
    def launch_count(self):
        return self._lib.hmi_launch_count(self._h)

    @property
    def kernel_name(self):
        return self._lib.hmi_kernel_name(self._h).decode()


class Group:
    """One grid cut into row slabs over several GPUs of this process (hmi_group_*): the engine behind
    ``Inpainter(..., devices=[...])``.  Same sweeps()/sweeps_norm()/run() surface as Solver."""

    def __init__(self, params, devices, exchange_every=0):
        self._lib = C.lib()
        self.params = params
        devs = (ctypes.c_int32 * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        C.check(self._lib.hmi_group_create(ctypes.byref(params), devs, len(devices), int(exchange_every), ctypes.byref(h)))
        self._h = h
        self.np_dtype = np.float64 if params.dtype == C.HMI_F64 else np.float32
        n, e, g = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        C.check(self._lib.hmi_group_info(self._h, ctypes.byref(n), ctypes.byref(e), ctypes.byref(g)))
        self.n_slabs, self.exchange_every, self.ghost_rows = n.value, e.value, g.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hmi_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_problem_host(self, image, mask):
        image = np.ascontiguousarray(image, dtype=self.np_dtype)
        m8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == np.bool_ else \
            np.ascontiguousarray(mask, dtype=np.uint8)
        want = (self.params.rows, self.params.cols)
        if image.shape != want or m8.shape != want:
            raise ValueError("image/mask shape %s/%s, group expects %s" % (image.shape, m8.shape, want))
        C.check(self._lib.hmi_group_set_problem_host(self._h, image.ctypes.data, image.strides[0],
                                                     m8.ctypes.data, m8.strides[0]))

    def fill_unknown(self, value):
        C.check(self._lib.hmi_group_fill_unknown(self._h, float(value)))

    def set_problem_pyramid(self, finest, stride, coarse, xtab, ytab, scrub_nan=False):
        """Multigrid: make this group pyramid level log2(stride) of the problem `finest` holds — image and mask
        gathered from it on the device, unknown cells <- the up-sampled current iterate of `coarse`
        (hmi_group_set_problem_pyramid).  stride 1: `finest` itself, in place.  xtab / ytab are the
        (src0, src1, w0, w1) tables of cv2's INTER_LINEAR per axis.  Returns HmiBlendInfo."""
        p = self.params
        tabs = []
        for (s0, s1, w0, w1), n in ((xtab, p.cols), (ytab, p.rows)):
            tabs += [np.ascontiguousarray(s0, dtype=np.int32), np.ascontiguousarray(s1, dtype=np.int32),
                     np.ascontiguousarray(w0, dtype=self.np_dtype), np.ascontiguousarray(w1, dtype=self.np_dtype)]
            if any(t.shape != (n,) for t in tabs[-4:]):
                raise ValueError("interpolation tables must have one entry per output column / row")
        info = C.HmiBlendInfo()
        C.check(self._lib.hmi_group_set_problem_pyramid(self._h, finest._h, int(stride), coarse._h,
                                                        *[t.ctypes.data for t in tabs], 1 if scrub_nan else 0,
                                                        ctypes.byref(info)))
        return info

    def sweeps(self, n):
        C.check(self._lib.hmi_group_sweeps(self._h, int(n)))

    def sweeps_norm(self, n):
        nm = C.HmiNorm()
        C.check(self._lib.hmi_group_sweeps_norm(self._h, int(n), ctypes.byref(nm)))
        return Norm(nm.sum_d2, nm.sum_f2, nm.max_abs, nm.has_nan)

    def run(self):
        st = C.HmiStatus()
        C.check(self._lib.hmi_group_run(self._h, ctypes.byref(st)))
        return st

    def sync(self):
        C.check(self._lib.hmi_group_sync(self._h))

    def set_term_thres(self, term_thres):
        C.check(self._lib.hmi_group_set_term_thres(self._h, float(term_thres)))
        self.params.term_thres = float(term_thres)

    def get_result_host(self, out=None):
        p = self.params
        if out is None:
            out = np.empty((p.rows, p.cols), dtype=self.np_dtype)
            if out.nbytes >= (1 << 22):
                self._lib.hmi_host_prefault(out.ctypes.data, out.nbytes)
        C.check(self._lib.hmi_group_get_result_host(self._h, out.ctypes.data, out.strides[0]))
        return out

    def slab(self, i):
        """(raw solver handle, first row, end row) of slab i — for tests."""
        h, a, b = ctypes.c_void_p(), ctypes.c_int32(), ctypes.c_int32()
        C.check(self._lib.hmi_group_slab(self._h, int(i), ctypes.byref(h), ctypes.byref(a), ctypes.byref(b)))
        return h, a.value, b.value

    @property
    def launch_count(self):
        return self._lib.hmi_group_launch_count(self._h)

    @property
    def kernel_name(self):
        h, _, _ = self.slab(0)
        return self._lib.hmi_kernel_name(h).decode()

    @property
    def temporal_block(self):
        h, _, _ = self.slab(0)
        return self._lib.hmi_temporal_block(h)


class AreaSession:
    """The elevation grid, mask_int and elevation_int resident on one GPU for the reference CLI's loop over
    work areas (hmi_areas_*): windows are cropped, masked, NaN-scrubbed and pasted back on the device."""

    def __init__(self, elevation, mask_int, device=0):
        self._lib = C.lib()
        self.elevation = np.ascontiguousarray(elevation)
        if self.elevation.dtype not in (np.float32, np.float64) or self.elevation.ndim != 2:
            raise TypeError("elevation must be a 2-D float32 or float64 array")
        self.np_dtype = self.elevation.dtype.type
        m8 = np.ascontiguousarray(mask_int, dtype=bool).view(np.uint8)
        if m8.shape != self.elevation.shape:
            raise ValueError("elevation and mask_int must have the same shape")
        h = ctypes.c_void_p()
        C.check(self._lib.hmi_areas_create(self.elevation.shape[0], self.elevation.shape[1], dtype_code(self.elevation.dtype),
                                           int(device), self.elevation.ctypes.data, self.elevation.strides[0],
                                           m8.ctypes.data, m8.strides[0], ctypes.byref(h)))
        self._h = h
        self.shape = self.elevation.shape

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hmi_areas_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_work_area(self, area):
        """`area`: bool layer of the grid's shape, or None for one area covering the grid."""
        if area is None:
            C.check(self._lib.hmi_areas_set_work_area(self._h, None, 0))
            return
        a8 = np.ascontiguousarray(area, dtype=bool).view(np.uint8)
        if a8.shape != self.shape:
            raise ValueError("work area must have the grid's shape")
        C.check(self._lib.hmi_areas_set_work_area(self._h, a8.ctypes.data, a8.strides[0]))

    def load_window(self, group, r0, c0):
        info = C.HmiWindowInfo()
        C.check(self._lib.hmi_areas_load_window(self._h, group._h, int(r0), int(c0), ctypes.byref(info)))
        return info

    def paste(self, group, r0, c0):
        C.check(self._lib.hmi_areas_paste(self._h, group._h, int(r0), int(c0)))

    def result(self):
        out = np.empty(self.shape, dtype=self.np_dtype)
        if out.nbytes >= (1 << 22):
            self._lib.hmi_host_prefault(out.ctypes.data, out.nbytes)
        C.check(self._lib.hmi_areas_get_result_host(self._h, out.ctypes.data, out.strides[0]))
        return out


def inpaint_host_devices(params, devices, image, mask, out=None, fill=None, init_host=False):
    """hmi_inpaint_host_devices: the one-shot call over several GPUs of this process."""
    lib = C.lib()
    np_dtype = np.float64 if params.dtype == C.HMI_F64 else np.float32
    image = np.ascontiguousarray(image, dtype=np_dtype)
    m8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == np.bool_ else \
        np.ascontiguousarray(mask, dtype=np.uint8)
    if out is None:
        out = np.empty_like(image)
        if out.nbytes >= (1 << 22):
            lib.hmi_host_prefault(out.ctypes.data, out.nbytes)
    elif out.shape != image.shape or out.dtype != image.dtype or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous array of the image's shape and dtype")
    devs = (ctypes.c_int32 * len(devices))(*[int(d) for d in devices])
    st = C.HmiStatus()
    C.check(lib.hmi_inpaint_host_devices(ctypes.byref(params), devs, len(devices), image.ctypes.data, m8.ctypes.data,
                                         0 if fill is None else (2 if init_host else 1),
                                         0.0 if fill is None else float(fill),
                                         out.ctypes.data, ctypes.byref(st)))
    return out, st


def inpaint_host(params, image, mask, out=None, fill=None, init_host=False):
    """hmi_inpaint_host: one call, host buffers in, host buffer out.  `out` (optional) is a
    C-contiguous array of the grid's shape and dtype to receive the result, e.g. pinned memory.
    `fill` (optional) = value the unknown cells take on the device after the upload
    (hmi_inpaint_host_fill: the 'zeros' / 'mean' initialiser fused in); with `init_host` the unknown
    cells of `image` itself take it too, once the upload has read the array (hmi_inpaint_host_init)."""
    lib = C.lib()
    np_dtype = np.float64 if params.dtype == C.HMI_F64 else np.float32
    image = np.ascontiguousarray(image, dtype=np_dtype)
    m8 = np.ascontiguousarray(mask).view(np.uint8) if mask.dtype == np.bool_ else \
        np.ascontiguousarray(mask, dtype=np.uint8)
    if out is None:
        out = np.empty_like(image)
        if out.nbytes >= (1 << 22):
            lib.hmi_host_prefault(out.ctypes.data, out.nbytes)
    elif out.shape != image.shape or out.dtype != image.dtype or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous array of the image's shape and dtype")
    st = C.HmiStatus()
    if fill is None:
        C.check(lib.hmi_inpaint_host(ctypes.byref(params), image.ctypes.data, m8.ctypes.data,
                                     out.ctypes.data, ctypes.byref(st)))
    else:
        fn = lib.hmi_inpaint_host_init if init_host else lib.hmi_inpaint_host_fill
        C.check(fn(ctypes.byref(params), image.ctypes.data, m8.ctypes.data, float(fill), out.ctypes.data,
                   ctypes.byref(st)))
    return out, st
