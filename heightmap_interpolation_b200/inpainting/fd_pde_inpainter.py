"""B200 drop-in for the reference's FD-PDE inpainter base class.

Mirrors ``heightmap_interpolation/inpainting/fd_pde_inpainter.py`` of the reference (v1.0.7):
same constructor keywords, defaults and ``ValueError`` messages (:71-100), ``get_config`` (:124-139),
``set_term_criteria`` (:141-158), ``inpaint`` (:160-185) and the loop semantics of ``inpaint_grid``
(:281-366).  The loop body itself — step function, mask-constrained update, over-relaxation and the
convergence norm — runs in hand-written CUDA kernels behind the C ABI of ``include/hmi_b200.h``.

Extra keywords (ignored by the reference, which silently drops unknown keys):
``device`` (CUDA ordinal), ``devices`` (list of ordinals or "all": the grid is cut into row slabs over
these GPUs of this process, include/hmi_b200.h hmi_group_*), ``kernel`` ("auto" | "generic" | "stream"),
``temporal_block`` (sweeps fused per launch, 0 = auto), ``nearest_backend`` ("b200" | "scipy": who
evaluates ``init_with="nearest"``).  After a solve the object exposes ``iterations_``, ``last_diff_`` and
``converged_`` (the reference only prints them).
"""
from abc import ABC
from timeit import default_timer as timer

import numpy as np

from .. import _capi as C
from ..solver import Group, Norm, Solver, inpaint_host, inpaint_host_devices, make_params
from .convolver import Convolver
from .initializer import Initializer


class FDPDEInpainter(ABC):
    """Abstract base of the explicit finite-difference PDE inpainters (reference class of the same name)."""

    METHOD = None  # hmi_method code, set by subclasses
    # below this size the host fill (~0.14 ms per million cells) is cheaper than starting a thread for it
    FUSE_INIT_MIN_CELLS = 1 << 21

    def __init__(self, **kwargs):
        super().__init__()
        # --- same keywords and defaults as the reference (fd_pde_inpainter.py:71-84)
        self.dt = kwargs.pop("update_step_size", 0.01)
        self.term_check_iters = kwargs.pop("term_check_iters", 1000)
        self.term_thres = kwargs.pop("term_thres", 1e-8)
        self.max_iters = int(kwargs.pop("max_iters", 1e8))
        self.relaxation = kwargs.pop("relaxation", 0)
        self.print_progress = kwargs.pop("print_progress", False)
        self.print_progress_iters = kwargs.pop("print_progress_iters", 1000)
        self.mgs_levels = kwargs.pop("mgs_levels", 1)
        self.mgs_min_res = kwargs.pop("mgs_min_res", 100)
        self.init_with = kwargs.pop("init_with", "zeros")
        self.convolver_type = kwargs.pop("convolver", "masked")
        self.convolver = Convolver(self.convolver_type)
        self.debug_dir = kwargs.pop("debug_dir", "")
        self.term_criteria = kwargs.pop("term_criteria", "relative")
        self.ts = 0
        # --- B200 back-end keywords
        self.device = kwargs.pop("device", None)
        self.devices = kwargs.pop("devices", None)
        self.convolver.device = 0 if self.device is None else int(self.device)
        self.kernel = kwargs.pop("kernel", "auto")
        self.temporal_block = int(kwargs.pop("temporal_block", 0))

        # --- same validation, same messages (fd_pde_inpainter.py:87-100)
        if self.dt <= 0:
            raise ValueError("update_step_size must be larger than zero")
        if self.term_thres <= 0:
            raise ValueError("term_thres must be larger than zero")
        if self.max_iters <= 0:
            raise ValueError("max_iters must be larger than zero")
        if self.relaxation != 0.0 and (self.relaxation < 1.0 or self.relaxation > 2.0):
            raise ValueError("relaxation must be a number between 1 and 2 (0 to deactivate)")
        if not isinstance(self.max_iters, int):
            raise ValueError("max_iters must be an integer")
        if not isinstance(self.mgs_levels, int):
            raise ValueError("mgs_levels must be an integer")
        if not isinstance(self.mgs_min_res, int):
            raise ValueError("mgs_min_res must be an integer")
        if self.kernel not in C.KERNELS:
            raise ValueError("kernel must be one of: " + ", ".join(C.KERNELS))
        if self.debug_dir:
            raise NotImplementedError("debug_dir (matplotlib PNG dumps) is not part of the B200 back-end")

        self.map_term_criteria_str_to_int = {"relative": 0, "absolute": 1, "absolute_percent": 2}
        self.print_progress_table_row_str = "|{:>11d}|{:>17.10f}|"
        self.print_progress_last_table_row_str = "| CONVERGED |{:>17.10f}|"
        self.initializer = Initializer(self.init_with, nearest_backend=kwargs.pop("nearest_backend", "b200"),
                                       device=self.device)

        # results of the last solve
        self.iterations_ = None
        self.last_diff_ = None
        self.converged_ = None
        self.kernel_name_ = None
        self.launches_ = None
        self.history_ = []          # (rows, cols, iterations) of every inpaint_grid call of the last inpaint()

    # ------------------------------------------------------------------ configuration
    def get_config(self):
        return {"update_step_size": self.dt,
                "term_criteria": self.term_criteria,
                "term_check_iters": self.term_check_iters,
                "term_thres": self.term_thres,
                "max_iters": self.max_iters,
                "relaxation": self.relaxation,
                "print_progress": self.print_progress,
                "print_progress_iters": self.print_progress_iters,
                "mgs_levels": self.mgs_levels,
                "mgs_min_res": self.mgs_min_res,
                "init_with": self.init_with,
                "convolver": self.convolver_type}

    def set_term_criteria(self, f):
        """What fd_pde_inpainter.py:141-158 does: resolve the criterion name and, for 'absolute_percent',
        turn the percentage into an absolute threshold using the z-range of the initialised grid `f` —
        in place on the object, so a second solve with the same inpainter compounds it."""
        kind = self.map_term_criteria_str_to_int.get(self.term_criteria, -1)
        self.term_criteria_int = kind
        if kind not in (0, 1, 2):
            raise ValueError("Unknown termination criteria!")
        if kind < 2:
            self.print_msg("* Termination criteria --> {:s} change = {:f}".format(("relative", "absolute")[kind],
                                                                                 self.term_thres))
            return
        z_range = abs(np.max(f) - np.min(f))
        percent = self.term_thres
        self.term_thres = z_range * percent
        for line in ("* Termination criteria --> absolute percent change:",
                     f"    - Abs. Z range: {z_range:f}",
                     f"    - Percent = {percent:f}",
                     f"    - Terminate if absolute change < {self.term_thres:f}"):
            self.print_msg(line)

    # subclass hooks ----------------------------------------------------------------
    def _method_params(self):
        """Extra numeric options of the subclass (tension / epsilon)."""
        return {}

    def step_fun(self, f, mask=None):
        """One evaluation of the subclass's step function on the GPU (hmi_step_fun_host): the reference's
        hook (:406-408).  `mask` is the work mask the reference passes to its convolvers (True = evaluate;
        None = everywhere); the convolvers that ignore it (opencv, scipy-*) ignore it here too."""
        import ctypes
        from .convolver import _MASKED
        f = np.ascontiguousarray(f)
        if f.ndim != 2 or f.dtype not in (np.float32, np.float64):
            raise TypeError("f must be a 2-D float32 or float64 array")
        m8 = None
        if mask is not None and self.convolver.method in _MASKED and not getattr(self, "convolve_in_1d", False):
            m8 = np.ascontiguousarray(mask, dtype=np.uint8)
            if m8.shape != f.shape:
                raise ValueError("mask must have the same shape as f")
        self.term_criteria_int = self.map_term_criteria_str_to_int.get(self.term_criteria, 0)
        params = self._make_params(f.shape[0], f.shape[1], f.dtype)
        out = np.empty_like(f)
        C.check(C.lib().hmi_step_fun_host(ctypes.byref(params), f.ctypes.data, f.strides[0],
                                          None if m8 is None else m8.ctypes.data, 0 if m8 is None else m8.strides[0],
                                          out.ctypes.data, out.strides[0]))
        return out

    # ------------------------------------------------------------------ public API
    def inpaint(self, image, mask, out=None):
        """Inpaint `image` where `mask` is False (True == known).  Same contract as the reference
        (:160-185): the caller's `image` is initialised IN PLACE, a new array is returned.
        `out` is an extension: a preallocated (e.g. pinned) array to receive the result."""
        self.print_msg("*** INPAINTING ***")
        self.history_ = []
        if self.mgs_levels > 1:
            self.print_msg("* Using a multi-grid solver")
            from .multigrid import inpaint_multigrid
            return inpaint_multigrid(self, image, mask)
        self.print_msg("* Initializing the inpainting problem using the '{:s}' filler".format(self.init_with))
        if self._can_fuse_init(image, mask):
            return self._inpaint_fused_init(image, mask, out)
        image = self.initializer.initialize(image, mask)
        self.print_msg("* Optimization:")
        return self.inpaint_grid(image, mask, out)

    # -- 'zeros' / 'mean' fused into the upload: the device fills its copy while a host thread
    #    initialises the caller's array in place (what the reference does first, initializer.py:15-18)
    def _can_fuse_init(self, image, mask):
        return (self.init_with.lower() in ("zeros", "mean") and not self.print_progress
                and isinstance(image, np.ndarray) and isinstance(mask, np.ndarray) and image.ndim == 2
                and image.dtype in (np.float32, np.float64) and mask.dtype == np.bool_
                and mask.shape == image.shape and image.flags.c_contiguous and image.flags.writeable
                and mask.flags.c_contiguous and image.size >= self.FUSE_INIT_MIN_CELLS and self.term_check_iters > 0)

    def _inpaint_fused_init(self, image, mask, out):
        import ctypes
        lib = C.lib()
        code = C.HMI_F64 if image.dtype == np.float64 else C.HMI_F32
        if self.init_with.lower() == "zeros":
            value = 0.0
        else:
            m = ctypes.c_double()
            C.check(lib.hmi_host_known_mean(image.ctypes.data, mask.ctypes.data, image.size, code, ctypes.byref(m)))
            value = float(image.dtype.type(m.value))
        if self.term_criteria == "absolute_percent":
            # z-range of the *initialised* grid (fd_pde_inpainter.py:149-155) = known cells + fill value
            lo, hi, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            C.check(lib.hmi_host_known_minmax(image.ctypes.data, mask.ctypes.data, image.size, code,
                                              ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(cnt)))
            lo, hi = lo.value, hi.value
            if cnt.value < image.size and not (lo != lo):
                lo, hi = (min(lo, value), max(hi, value)) if cnt.value else (value, value)
            self.set_term_criteria(_RangeView(image.dtype.type(lo), image.dtype.type(hi)))
        else:
            self.set_term_criteria(image)
        self.print_msg("* Optimization:")
        params = self._make_params(image.shape[0], image.shape[1], image.dtype)
        # the library initialises the caller's array in place once the upload has read it
        out, st = self._solve_host(params, image, mask, out, fill=value, init_host=True)
        self._record(st.updates_done, st.last_diff, bool(st.converged), None, None)
        self.history_.append((image.shape[0], image.shape[1], self.iterations_))
        if not self.converged_:
            print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")
        return out

    def _device(self):
        if self.device is None:
            devs = self._device_list()
            return devs[0] if devs else 0
        return int(self.device)

    def _device_list(self):
        """The `devices` keyword resolved to a list of CUDA ordinals, or None for a single-GPU solve."""
        if self.devices is None:
            return None
        if isinstance(self.devices, str):
            if self.devices != "all":
                raise ValueError("devices must be a list of CUDA ordinals or 'all'")
            n = C.check(C.lib().hmi_device_count())
            devs = list(range(n))
        else:
            devs = [int(d) for d in self.devices]
        if not devs:
            raise ValueError("devices must name at least one GPU")
        return devs if len(devs) > 1 else None

    def _make_params(self, rows, cols, dtype, **over):
        kw = dict(rows=rows, cols=cols, dtype=dtype, method=self.METHOD,
                  orientation=self.convolver.orientation, border=self.convolver.border,
                  criteria=C.HMI_RELATIVE if self.term_criteria_int == 0 else C.HMI_ABSOLUTE,
                  dt=self.dt, relaxation=self.relaxation, term_thres=self.term_thres,
                  term_check_iters=self.term_check_iters, max_iters=self.max_iters,
                  device=self._device(), kernel=C.KERNELS[self.kernel],
                  temporal_block=self.temporal_block)
        kw.update(self._method_params())
        kw.update(over)
        return make_params(**kw)

    def inpaint_grid(self, image, mask, out=None):
        """Single-scale solve (fd_pde_inpainter.py:281-366) on one B200."""
        image, mask = _check_inputs(image, mask)
        self.set_term_criteria(image)
        if self.term_check_iters <= 0:
            raise ZeroDivisionError("integer modulo by zero")  # what `i % 0` raises in the reference (:322)
        params = self._make_params(image.shape[0], image.shape[1], image.dtype)
        if not self.print_progress:
            out, st = self._solve_host(params, image, mask, out)
            self._record(st.updates_done, st.last_diff, bool(st.converged), None, None)
        else:
            devs = self._device_list()
            with (Group(params, devs) if devs else Solver(params)) as s:
                s.set_problem_host(image, mask)
                done, diff, conv = self._loop(s, params.criteria)
                out = s.get_result_host(out)
                self._record(done, diff, conv, s.kernel_name, s.launch_count)
        self.history_.append((image.shape[0], image.shape[1], self.iterations_))
        if not self.converged_:
            print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")
        return out

    def _solve_host(self, params, image, mask, out, fill=None, init_host=False):
        """One-shot solve over host buffers on one GPU or, with `devices`, on several of this process."""
        devs = self._device_list()
        if devs:
            return inpaint_host_devices(params, devs, image, mask, out, fill=fill, init_host=init_host)
        return inpaint_host(params, image, mask, out, fill=fill, init_host=init_host)

    # -- the CLI's per-area solve with the grid resident on the device (apps/interpolate.py)
    def can_inpaint_window(self):
        """Initialisers that exist on the device (the others need the host: initializer.py:19-41)."""
        return (self.init_with.lower() in ("zeros", "mean") and self.mgs_levels <= 1 and self.term_check_iters > 0)

    def inpaint_window(self, session, r0, c0, rows, cols):
        """inpaint() of the window [r0, r0 + rows) x [c0, c0 + cols) of an AreaSession: crop, known mask, NaN -> 0,
        initialiser, solve and paste of the inpainted cells all happen on the device.  Same loop semantics and
        thresholds as inpaint() on the cropped arrays (the 'mean' fill value is accumulated in double)."""
        self.print_msg("*** INPAINTING ***")
        self.history_ = []
        self.print_msg("* Initializing the inpainting problem using the '{:s}' filler".format(self.init_with))
        self.term_criteria_int = self.map_term_criteria_str_to_int.get(self.term_criteria, -1)
        if self.term_criteria_int < 0:
            raise ValueError("Unknown termination criteria!")
        dtype = session.np_dtype
        params = self._make_params(rows, cols, dtype, term_thres=self.term_thres if self.term_thres > 0 else 1.0)
        devs = self._device_list() or [self._device()]
        with Group(params, devs) as g:
            info = session.load_window(g, r0, c0)
            if self.init_with.lower() == "zeros":
                value = 0.0
            else:
                value = float(dtype(info.known_sum / info.known)) if info.known else float("nan")
            lo, hi = info.known_min, info.known_max
            if info.unknown and value == value:
                lo, hi = (min(lo, value), max(hi, value)) if info.known else (value, value)
            self.set_term_criteria(_RangeView(dtype(lo), dtype(hi)))
            g.set_term_thres(self.term_thres)
            params.term_thres = float(self.term_thres)
            g.fill_unknown(value)
            self.print_msg("* Optimization:")
            if self.print_progress:
                done, diff, conv = self._loop(g, params.criteria)
            else:
                st = g.run()
                done, diff, conv = st.updates_done, st.last_diff, bool(st.converged)
            self._record(done, diff, conv, g.kernel_name, g.launch_count)
            self.history_.append((rows, cols, self.iterations_))
            if not self.converged_:
                print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")
            session.paste(g, r0, c0)

    def _record(self, done, diff, conv, kernel_name, launches):
        self.iterations_, self.last_diff_, self.converged_ = int(done), float(diff), bool(conv)
        self.kernel_name_, self.launches_ = kernel_name, launches

    def _loop(self, engine, criteria):
        """The reference loop, event driven: sweeps are issued in batches up to the next iteration
        index at which the reference does something observable (a check at i % T == 0, a progress
        row at i % P == 0, or the last iteration).  `engine` provides sweeps(n) / sweeps_norm(n)."""
        T, P, maxit = int(self.term_check_iters), int(self.print_progress_iters), int(self.max_iters)
        show = self.print_progress and P > 0
        diff = 100000
        terminate = False
        i = 0
        while i < maxit:
            nxt = (i + T - 1) // T * T
            if show:
                nxt = min(nxt, (i + P - 1) // P * P)
            e = min(nxt, maxit - 1)
            n = e - i + 1
            is_check = (e % T == 0)
            if is_check:
                nm = engine.sweeps_norm(n)
                diff = nm.diff(criteria)
                terminate = diff < self.term_thres
            else:
                engine.sweeps(n)
            if show and e % P == 0:
                if e == 0:
                    print("+-----------+-----------------+")
                    print("| Iteration | Function change |")
                    print("+-----------+-----------------+")
                print(self.print_progress_table_row_str.format(e, diff))
            i = e + 1
            if is_check and terminate:
                if self.print_progress:
                    print("+-----------+-----------------+")
                    print(self.print_progress_last_table_row_str.format(diff))
                    print("+-----------+-----------------+")
                return i, diff, True
        return i, diff, False

    # ------------------------------------------------------------------ verbose output (reference :386-404)
    def print_msg(self, msg):
        if self.print_progress:
            print(msg)

    def print_start(self, msg):
        """`msg` without a newline and start the stopwatch; print_end() closes the line with the time."""
        if not self.print_progress:
            return
        print(msg, end="")
        self.ts = timer()

    def print_end(self):
        if self.print_progress:
            print("done ({:.2f} s)".format(timer() - self.ts))


class _RangeView:
    """Stands in for a grid in set_term_criteria(): only max() and min() are taken."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def max(self, *a, **k):
        return self.hi

    def min(self, *a, **k):
        return self.lo


def _check_inputs(image, mask):
    image = np.asarray(image)
    mask = np.asarray(mask)
    if image.ndim != 2:
        raise ValueError("image must be a 2-D array")
    if mask.shape != image.shape:
        raise ValueError("mask must have the same shape as the image")
    if image.dtype not in (np.float32, np.float64):
        raise TypeError("image dtype must be float32 or float64, got %s" % image.dtype)
    if mask.dtype != np.bool_:
        mask = mask != 0
    return np.ascontiguousarray(image), np.ascontiguousarray(mask)
