"""ctypes binding of the C ABI declared in include/hmi_b200.h.

The CUDA library is the product: if ``csrc/libhmi_b200.so`` is missing or cannot be loaded this
module raises — there is no CPU or PyTorch fallback behind it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HMI_B200_LIB") or os.path.join(_HERE, "csrc", "libhmi_b200.so")

HMI_F32, HMI_F64 = 0, 1
HMI_HARMONIC, HMI_CCST, HMI_TV, HMI_AMLE = 0, 1, 2, 3
HMI_BORDER_REFLECT101, HMI_BORDER_SYMMETRIC, HMI_BORDER_ZERO = 0, 1, 2
HMI_RELATIVE, HMI_ABSOLUTE = 0, 1
HMI_KERNEL_AUTO, HMI_KERNEL_GENERIC, HMI_KERNEL_STREAM = 0, 1, 2
HMI_ERR_BAD_ARG, HMI_ERR_CUDA, HMI_ERR_NO_DEVICE, HMI_ERR_STATE, HMI_ERR_UNSUPPORTED = -1, -2, -3, -4, -5

KERNELS = {"auto": HMI_KERNEL_AUTO, "generic": HMI_KERNEL_GENERIC, "stream": HMI_KERNEL_STREAM}


class HmiParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("device", ctypes.c_int32),
        ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
        ("dtype", ctypes.c_int32), ("method", ctypes.c_int32),
        ("orientation", ctypes.c_int32), ("border", ctypes.c_int32),
        ("criteria", ctypes.c_int32), ("kernel", ctypes.c_int32),
        ("temporal_block", ctypes.c_int32), ("ghost_top", ctypes.c_int32),
        ("ghost_bottom", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("dt", ctypes.c_double), ("tension", ctypes.c_double), ("epsilon", ctypes.c_double),
        ("relaxation", ctypes.c_double), ("term_thres", ctypes.c_double),
        ("term_check_iters", ctypes.c_int64), ("max_iters", ctypes.c_int64),
    ]


class HmiStatus(ctypes.Structure):
    _fields_ = [("updates_done", ctypes.c_int64), ("last_diff", ctypes.c_double),
                ("converged", ctypes.c_int32), ("checks", ctypes.c_int32)]


class HmiBlendInfo(ctypes.Structure):
    _fields_ = [("min", ctypes.c_double), ("max", ctypes.c_double),
                ("has_nan", ctypes.c_int32), ("image_had_nan", ctypes.c_int32)]


class HmiHaloInfo(ctypes.Structure):
    _fields_ = [("block", ctypes.c_void_p), ("block_bytes", ctypes.c_uint64), ("device", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("ipc_handle", ctypes.c_ubyte * 64)]


class HmiWindowInfo(ctypes.Structure):
    _fields_ = [("known", ctypes.c_int64), ("unknown", ctypes.c_int64), ("known_sum", ctypes.c_double),
                ("known_min", ctypes.c_double), ("known_max", ctypes.c_double)]


class HmiNorm(ctypes.Structure):
    _fields_ = [("sum_d2", ctypes.c_double), ("sum_f2", ctypes.c_double),
                ("max_abs", ctypes.c_double), ("has_nan", ctypes.c_double)]


# every symbol include/hmi_b200.h declares: name -> (restype, argtypes)
_vp, _sz, _i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64
SYMBOLS = {
    "hmi_abi_version": (ctypes.c_int, []),
    "hmi_last_error": (ctypes.c_char_p, []),
    "hmi_device_count": (ctypes.c_int, []),
    "hmi_create": (ctypes.c_int, [ctypes.POINTER(HmiParams), ctypes.POINTER(_vp)]),
    "hmi_destroy": (None, [_vp]),
    "hmi_set_problem_host": (ctypes.c_int, [_vp, _vp, _sz, _vp, _sz]),
    "hmi_set_problem_device": (ctypes.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "hmi_sweeps": (ctypes.c_int, [_vp, _i64, _vp]),
    "hmi_sweeps_overlap": (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    "hmi_sweeps_norm": (ctypes.c_int, [_vp, _i64, _vp, ctypes.POINTER(HmiNorm)]),
    "hmi_halo_setup": (ctypes.c_int, [_vp, ctypes.POINTER(HmiHaloInfo)]),
    "hmi_halo_connect": (ctypes.c_int, [_vp, ctypes.c_int32, _vp, ctypes.c_int32]),
    "hmi_halo_connect_ipc": (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.c_char_p]),
    "hmi_sweeps_exchange": (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    "hmi_group_create": (ctypes.c_int, [ctypes.POINTER(HmiParams), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                        ctypes.c_int32, ctypes.POINTER(_vp)]),
    "hmi_group_destroy": (None, [_vp]),
    "hmi_group_set_problem_host": (ctypes.c_int, [_vp, _vp, _sz, _vp, _sz]),
    "hmi_group_fill_unknown": (ctypes.c_int, [_vp, ctypes.c_double]),
    "hmi_group_sweeps": (ctypes.c_int, [_vp, _i64]),
    "hmi_group_sweeps_norm": (ctypes.c_int, [_vp, _i64, ctypes.POINTER(HmiNorm)]),
    "hmi_group_run": (ctypes.c_int, [_vp, ctypes.POINTER(HmiStatus)]),
    "hmi_group_set_term_thres": (ctypes.c_int, [_vp, ctypes.c_double]),
    "hmi_group_get_result_host": (ctypes.c_int, [_vp, _vp, _sz]),
    "hmi_group_sync": (ctypes.c_int, [_vp]),
    "hmi_group_set_problem_pyramid": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                     ctypes.c_int32, ctypes.POINTER(HmiBlendInfo)]),
    "hmi_group_info": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int32)]),
    "hmi_group_slab": (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int32)]),
    "hmi_group_launch_count": (_i64, [_vp]),
    "hmi_inpaint_host_devices": (ctypes.c_int, [ctypes.POINTER(HmiParams), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                                _vp, _vp, ctypes.c_int32, ctypes.c_double, _vp,
                                                ctypes.POINTER(HmiStatus)]),
    "hmi_run": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(HmiStatus)]),
    "hmi_get_result_host": (ctypes.c_int, [_vp, _vp, _sz]),
    "hmi_get_result_device": (ctypes.c_int, [_vp, _vp, _sz, _vp]),
    "hmi_inpaint_host": (ctypes.c_int, [ctypes.POINTER(HmiParams), _vp, _vp, _vp,
                                        ctypes.POINTER(HmiStatus)]),
    "hmi_inpaint_host_fill": (ctypes.c_int, [ctypes.POINTER(HmiParams), _vp, _vp, ctypes.c_double, _vp,
                                             ctypes.POINTER(HmiStatus)]),
    "hmi_inpaint_host_init": (ctypes.c_int, [ctypes.POINTER(HmiParams), _vp, _vp, ctypes.c_double, _vp,
                                             ctypes.POINTER(HmiStatus)]),
    "hmi_set_problem_host_blend": (ctypes.c_int, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                  ctypes.c_int32, ctypes.POINTER(HmiBlendInfo)]),
    "hmi_step_fun_host": (ctypes.c_int, [ctypes.POINTER(HmiParams), _vp, _sz, _vp, _sz, _vp, _sz]),
    "hmi_convolve3x3_host": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, _sz,
                                            ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.c_int32, _vp, _sz,
                                            _vp, _sz]),
    "hmi_areas_create": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, _sz, _vp, _sz,
                                        ctypes.POINTER(_vp)]),
    "hmi_areas_destroy": (None, [_vp]),
    "hmi_areas_set_work_area": (ctypes.c_int, [_vp, _vp, _sz]),
    "hmi_areas_load_window": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(HmiWindowInfo)]),
    "hmi_areas_paste": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32]),
    "hmi_areas_get_result_host": (ctypes.c_int, [_vp, _vp, _sz]),
    "hmi_current_grid": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_sz)]),
    "hmi_mask_grid": (ctypes.c_int, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_sz)]),
    "hmi_host_fill_unknown": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int32, ctypes.c_double]),
    "hmi_host_known_mean": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]),
    "hmi_init_nearest_host": (ctypes.c_int, [_vp, _sz, _vp, _sz, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_int32]),
    "hmi_host_known_minmax": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64)]),
    "hmi_host_prefault": (ctypes.c_int, [_vp, _sz]),
    "hmi_release_cache": (ctypes.c_int, []),
    "hmi_upload_pipeline_count": (ctypes.c_int64, []),
    "hmi_debug_band_steps": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    "hmi_debug_pick_chunks": (ctypes.c_int, [ctypes.c_int32] * 6 + [ctypes.c_double, ctypes.c_int32,
                                             ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "hmi_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, _i64]),
    "hmi_set_term_thres": (ctypes.c_int, [_vp, ctypes.c_double]),
    "hmi_temporal_block": (ctypes.c_int, [_vp]),
    "hmi_launch_count": (_i64, [_vp]),
    "hmi_kernel_name": (ctypes.c_char_p, [_vp]),
}

_lib = None


def lib():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "heightmap_interpolation_b200: %s is missing — build it with "
                "`python -m heightmap_interpolation_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.hmi_abi_version() != 1:
            raise RuntimeError("hmi_b200 ABI version mismatch")
        _lib = L
    return _lib


class HmiError(RuntimeError):
    pass


def check(rc):
    """Map a negative hmi_error to the exception the reference's Python layer would raise."""
    if rc >= 0:
        return rc
    msg = lib().hmi_last_error().decode("utf-8", "replace")
    if rc == HMI_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == HMI_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise HmiError("hmi_b200 error %d: %s" % (rc, msg))
