"""Multigrid driver (reference: FDPDEInpainter.inpaint_multigrid, inpainting/fd_pde_inpainter.py:187-279,
and halve_image, misc/image_proc.py:4-12).

Same control flow as the reference: a pyramid by stride-2 decimation of image and mask, the
initialiser applied to the coarsest level only (in place on a strided view, so it writes through to
the caller's array exactly like the reference), the coarsest level solved with
``term_thres = orig * 2**num_levels``, then for each finer level L the coarser solution is
up-sampled bilinearly (OpenCV ``cv2.resize`` default, half-pixel centres), blended into the unknown
cells and relaxed with ``term_thres = orig * 2**L``.

Every level is relaxed on the B200 and the solution never leaves the device between levels: the
bilinear up-sampling and the blend run in one kernel (``hmi_set_problem_host_blend``) from 1-D tap /
weight tables built here exactly as OpenCV does, which also returns the z-range ``absolute_percent``
needs.  What is left on the host per level is a contiguous copy of the decimated view and its upload.
``resize_bilinear`` is the NumPy restatement of that kernel (tests pin both against cv2).
"""
import math

import numpy as np


def halve_image(img):
    """misc/image_proc.py:4-12 — a strided view, not a copy."""
    dims = img.shape
    if len(dims) == 3:
        return img[0:dims[0]:2, 0:dims[1]:2, :]
    elif len(dims) == 2:
        return img[0:dims[0]:2, 0:dims[1]:2]
    raise ValueError("Number of dimensions of the input image not 2 nor 3")


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


def inpaint_multigrid(self, image, mask):
    """`self` is an FDPDEInpainter; mirrors fd_pde_inpainter.py:187-279 line by line."""
    if image.shape[0] < self.mgs_min_res or image.shape[1] < self.mgs_min_res:
        print("[WARNING] A multigrid solver was requested, but the size of the image is too small, "
              "defaulting to a single-scale inpainting")
        return self.inpaint_grid(image, mask)

    self.print_start("Creating the image pyramid... ")
    image_pyramid = [image]
    mask_pyramid = [mask]
    num_levels = self.mgs_levels
    for level in range(1, self.mgs_levels):
        width = math.ceil(image_pyramid[level - 1].shape[1] / 2)
        height = math.ceil(image_pyramid[level - 1].shape[0] / 2)
        if width < self.mgs_min_res or height < self.mgs_min_res:
            print("[WARNING] Stopping pyramid construction at level {:d}, image resolution would be too "
                  "small at this level (width or height < {:d})".format(level, self.mgs_min_res))
            num_levels = level
            break
        image_rs = halve_image(image_pyramid[level - 1])
        image_pyramid.append(image_rs)
        mask_rs = halve_image(np.asarray(mask_pyramid[level - 1], dtype="uint8"))
        mask_valid = ~np.isnan(image_rs)
        mask_rs = np.logical_and(mask_rs, mask_valid)
        mask_pyramid.append(mask_rs)
    self.print_end()

    self.print_start("[Pyramid Level {:d}] Initializing the deepest level... ".format(num_levels - 1))
    init_lower_scale = self.initializer.initialize(image_pyramid[num_levels - 1], mask_pyramid[num_levels - 1])
    self.print_end()
    self.print_start("[Pyramid Level {:d}] Inpainting...\\n".format(num_levels - 1))
    original_term_thres = self.term_thres
    self.term_thres = original_term_thres * 2 ** (num_levels)
    coarse_mask = mask_pyramid[num_levels - 1] > 0
    if num_levels == 1:
        inpainted = self.inpaint_grid(init_lower_scale, coarse_mask)
        self.print_end()
        return inpainted

    from ..solver import Solver
    from .fd_pde_inpainter import _RangeView, _check_inputs

    def solve(solver, params, shape):
        if self.print_progress:
            done, diff, conv = self._loop(solver, params.criteria)
        else:
            st = solver.run()
            done, diff, conv = st.updates_done, st.last_diff, bool(st.converged)
        self._record(done, diff, conv, solver.kernel_name, solver.launch_count)
        self.history_.append((shape[0], shape[1], self.iterations_))
        if not self.converged_:
            print("[WARNING] Inpainting did NOT converge: Maximum number of iterations reached...")

    coarse = None
    try:
        # coarsest level: plain inpaint_grid semantics, but the solution stays on the device
        img_c, m_c = _check_inputs(init_lower_scale, coarse_mask)
        self.set_term_criteria(img_c)
        params = self._make_params(img_c.shape[0], img_c.shape[1], img_c.dtype)
        coarse = Solver(params)
        coarse.set_problem_host(img_c, m_c)
        solve(coarse, params, img_c.shape)
        self.print_end()
        for level in range(num_levels - 2, -1, -1):
            self.print_start("[Pyramid Level {:d}] Inpainting...\\n".format(level))
            self.term_thres = original_term_thres * 2 ** level
            image = image_pyramid[level]
            mask = mask_pyramid[level]
            img_c, m_c = _check_inputs(image, mask)
            wt = np.float32 if img_c.dtype == np.float32 else np.float64
            crows, ccols = coarse.params.rows, coarse.params.cols
            xtab = _linear_coeffs(ccols, img_c.shape[1], wt)
            ytab = _linear_coeffs(crows, img_c.shape[0], wt)
            # provisional threshold; 'absolute_percent' rescales it once the blended grid's range is known
            params = self._make_params(img_c.shape[0], img_c.shape[1], img_c.dtype,
                                       term_thres=self.term_thres if self.term_thres > 0 else 1.0)
            fine = Solver(params)
            try:
                info = fine.set_problem_host_blend(img_c, m_c, coarse, xtab, ytab, scrub_nan=(level == 0))
            except Exception:
                fine.close()
                raise
            if level == 0 and info.image_had_nan:      # the reference zeroes them in the caller's array (:259-260)
                image[np.isnan(image)] = 0
            self.set_term_criteria(_RangeView(img_c.dtype.type(info.min), img_c.dtype.type(info.max)))
            fine.set_term_thres(self.term_thres)
            params.term_thres = float(self.term_thres)
            coarse.close()
            coarse = fine
            solve(coarse, params, img_c.shape)
            self.print_end()
        return coarse.get_result_host()
    finally:
        if coarse is not None:
            coarse.close()
