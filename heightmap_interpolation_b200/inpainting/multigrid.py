"""Multigrid driver (reference: FDPDEInpainter.inpaint_multigrid, inpainting/fd_pde_inpainter.py:187-279,
and halve_image, misc/image_proc.py:4-12).

What the reference does: a pyramid by stride-2 decimation of image and mask (each coarser mask ANDed
with "image is not NaN"), the initialiser applied to the coarsest level only — in place on a strided
view, so it writes through to the caller's array —, the coarsest level solved with
``term_thres = orig * 2**num_levels``, then for each finer level L the coarser solution up-sampled
bilinearly (OpenCV ``cv2.resize`` default, half-pixel centres), blended into the unknown cells and
relaxed with ``term_thres = orig * 2**L``.

How it is done here: the pyramid lives on the GPU(s).  Level l of the reference's pyramid is every
2**l-th cell of the caller's array, so only two things touch the host: the coarsest level (a strided
view, tiny, initialised in place exactly like the reference) and ONE upload of the full-resolution
image and mask.  Every other level is gathered from that device copy with stride 2**l by the kernel that
also up-samples the coarser solution and blends it in (``hmi_group_set_problem_pyramid``), which returns
the z-range ``absolute_percent`` needs too; solutions never leave the device between levels, and each
level is a device group, so ``devices=[...]`` spreads every level over the GPUs (rows of the coarser
level are read across NVLink through peer memory).  The 1-D tap / weight tables are built here exactly
as OpenCV does; ``resize_bilinear`` is the NumPy restatement of the kernel (tests pin both against cv2).

One documented deviation: a *known* cell that holds NaN and sits on the coarsest level's lattice is
filled by the initialiser through the strided view; the reference evaluated its "not NaN" masks before
that, this driver after (its device copy is uploaded once, after the write-through), so at intermediate
levels such a cell keeps the initialiser's value instead of the up-sampled one.  The reference CLI
zeroes NaNs before inpainting (apps/interpolate_netcdf4.py:192), so the case does not arise there.
"""
import numpy as np


def _linear_coeffs(ssize, dsize, wt):
    """Source index pairs and weights of OpenCV's INTER_LINEAR along one axis (half-pixel centres):
    fx = (d + 0.5) * (ssize / dsize) - 0.5, clamped at both ends; weights (1 - fx, fx) in the
    image's own precision (verified against cv2 4.13 impulse responses)."""
    scale = float(ssize) / float(dsize)
    d = np.arange(dsize, dtype=np.float64)
    fx = (d + 0.5) * scale - 0.5
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx
    lo = sx < 0
    fx[lo] = 0.0
    sx[lo] = 0
    hi = sx >= ssize - 1
    fx[hi] = 0.0
    sx[hi] = ssize - 1
    sx1 = np.minimum(sx + 1, ssize - 1)
    fx = fx.astype(wt)
    return sx, sx1, (wt(1.0) - fx).astype(wt), fx


def resize_bilinear(src, rows, cols):
    """cv2.resize(src, dsize=(cols, rows)) with the default INTER_LINEAR, for 2-D float arrays."""
    src = np.asarray(src)
    wt = np.float32 if src.dtype == np.float32 else np.float64
    sx, sx1, a0, a1 = _linear_coeffs(src.shape[1], cols, wt)
    sy, sy1, b0, b1 = _linear_coeffs(src.shape[0], rows, wt)
    s = src.astype(wt, copy=False)
    tmp = s[:, sx] * a0[None, :] + s[:, sx1] * a1[None, :]
    out = tmp[sy, :] * b0[:, None] + tmp[sy1, :] * b1[:, None]
    return out.astype(src.dtype, copy=False)


def pyramid_shapes(shape, levels, min_res):
    """Grid sizes of the pyramid, finest first: each level halves (rounding up) the one before; the
    pyramid stops before the first level narrower or lower than `min_res` (fd_pde_inpainter.py:207-215).
    Returns (shapes, level at which construction stopped early or None)."""
    shapes = [(int(shape[0]), int(shape[1]))]
    for level in range(1, levels):
        rows, cols = -(-shapes[-1][0] // 2), -(-shapes[-1][1] // 2)
        if cols < min_res or rows < min_res:
            return shapes, level
        shapes.append((rows, cols))
    return shapes, None


def inpaint_multigrid(self, image, mask):
    """`self` is an FDPDEInpainter.  Same observable behaviour as fd_pde_inpainter.py:187-279: messages,
    per-level thresholds and iteration counts, write-through of the coarsest initialisation into the
    caller's array, NaN scrub of level 0, the mutated ``term_thres``."""
    from ..solver import Group
    from .fd_pde_inpainter import _RangeView, _check_inputs

    if image.shape[0] < self.mgs_min_res or image.shape[1] < self.mgs_min_res:
        print("[WARNING] A multigrid solver was requested, but the size of the image is too small, "
              "defaulting to a single-scale inpainting")
        return self.inpaint_grid(image, mask)

    self.print_start("Creating the image pyramid... ")
    shapes, stopped_at = pyramid_shapes(image.shape, self.mgs_levels, self.mgs_min_res)
    if stopped_at is not None:
        print("[WARNING] Stopping pyramid construction at level {:d}, image resolution would be too "
              "small at this level (width or height < {:d})".format(stopped_at, self.mgs_min_res))
    num_levels = len(shapes)
    top = num_levels - 1
    step = 2 ** top
    # the coarsest level: every `step`-th cell — a view of the caller's array, like the reference's pyramid
    coarse_view = image[::step, ::step]
    coarse_mask = np.asarray(mask[::step, ::step], dtype=bool)
    if top > 0:
        coarse_mask = coarse_mask & ~np.isnan(coarse_view)
    self.print_end()

    self.print_start("[Pyramid Level {:d}] Initializing the deepest level... ".format(top))
    init_lower_scale = self.initializer.initialize(coarse_view, coarse_mask)     # in place: writes through
    self.print_end()
    self.print_start("[Pyramid Level {:d}] Inpainting...\n".format(top))
    original_term_thres = self.term_thres
    self.term_thres = original_term_thres * 2 ** num_levels
    if num_levels == 1:
        inpainted = self.inpaint_grid(init_lower_scale, coarse_mask)
        self.print_end()
        return inpainted

    devices = self._device_list() or [self._device()]

    def solve(group, params, shape):
        if self.print_progress:
            done, diff, conv = self._loop(group, params.criteria)
        else:
            st = group.run()
            done, diff, conv = st.updates_done, st.last_diff, bool(st.converged)
        self._record(done, diff, conv, group.kernel_name, group.launch_count)
        self.history_.append((shape[0], shape[1], self.iterations_))
        if not self.converged_:
            print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")

    finest = coarse = None
    try:
        # coarsest level: plain inpaint_grid semantics from the (tiny) host arrays
        img_c, m_c = _check_inputs(init_lower_scale, coarse_mask)
        self.set_term_criteria(img_c)
        params = self._make_params(img_c.shape[0], img_c.shape[1], img_c.dtype)
        coarse = Group(params, devices)
        coarse.set_problem_host(img_c, m_c)
        # the one upload of the full-resolution problem (after the write-through above); the other levels
        # are gathered from it on the device
        img0, m0 = _check_inputs(image, mask)
        finest = Group(self._make_params(img0.shape[0], img0.shape[1], img0.dtype), devices)
        finest.set_problem_host(img0, m0)
        solve(coarse, params, img_c.shape)
        self.print_end()
        wt = np.float32 if img0.dtype == np.float32 else np.float64
        for level in range(top - 1, -1, -1):
            self.print_start("[Pyramid Level {:d}] Inpainting...\n".format(level))
            self.term_thres = original_term_thres * 2 ** level
            rows, cols = shapes[level]
            xtab = _linear_coeffs(coarse.params.cols, cols, wt)
            ytab = _linear_coeffs(coarse.params.rows, rows, wt)
            # provisional threshold; 'absolute_percent' rescales it once the blended grid's range is known
            params = self._make_params(rows, cols, img0.dtype, term_thres=self.term_thres if self.term_thres > 0 else 1.0)
            lvl = finest if level == 0 else Group(params, devices)
            try:
                info = lvl.set_problem_pyramid(finest, 2 ** level, coarse, xtab, ytab, scrub_nan=(level == 0))
            except Exception:
                if lvl is not finest:
                    lvl.close()
                raise
            if level == 0 and info.image_had_nan:      # the reference zeroes them in the caller's array (:259-260)
                image[np.isnan(image)] = 0
            self.set_term_criteria(_RangeView(img0.dtype.type(info.min), img0.dtype.type(info.max)))
            lvl.set_term_thres(self.term_thres)
            params.term_thres = float(self.term_thres)
            coarse.close()
            coarse = lvl
            solve(coarse, params, (rows, cols))
            self.print_end()
        return coarse.get_result_host()
    finally:
        if coarse is not None:
            coarse.close()
        if finest is not None and finest is not coarse:
            finest.close()
