"""The gridded (inpainting) branch of the reference CLI's ``interpolate()``, on arrays.

Reference: heightmap_interpolation/apps/interpolate_netcdf4.py:176-212 (same code in
apps/interpolate_xyz.py:255-279).  The reference loads a NetCDF grid, derives ``mask_int`` (True ==
cell to interpolate) and the ``work_areas`` stack, and for each work area crops elevation and mask
to the area's bounding box, zeroes NaNs, builds a fresh inpainter from the CLI parameters, inpaints
and pastes only the inpainted cells back.  File I/O stays with the reference; this function is what
its loop body becomes with ``--backend b200``::

    elevation_int = inpaint_work_areas(elevation, mask_int, work_areas, params)

``params`` is the argparse namespace of ``interpolate_netcdf4 <method> ...``
(create_inpainter_from_params, apps_common.py:172-215).
"""
import numpy as np

from .apps_common import FD_PDE_METHODS, create_inpainter_from_params


def inpaint_work_areas(elevation, mask_int, work_areas=None, params=None, inpainter_factory=None):
    """Returns ``elevation_int``: a copy of ``elevation`` with the cells of ``mask_int`` (True == to be
    interpolated) that fall inside each work area replaced by the inpainted values.

    ``work_areas``: bool array [rows, cols, n_areas] (None = one area covering the grid, what the
    reference builds without ``--areas``, netcdf_data_io.py:262).  ``inpainter_factory`` overrides
    ``create_inpainter_from_params(params)`` (a fresh inpainter per area, like the reference).
    """
    elevation = np.asarray(elevation)
    mask_int = np.asarray(mask_int, dtype=bool)
    if elevation.ndim != 2 or mask_int.shape != elevation.shape:
        raise ValueError("elevation and mask_int must be 2-D arrays of the same shape")
    if work_areas is None:
        work_areas = np.ones(elevation.shape + (1,), dtype=bool)
    work_areas = np.asarray(work_areas, dtype=bool)
    if work_areas.ndim == 2:
        work_areas = work_areas[:, :, None]
    if work_areas.shape[:2] != elevation.shape:
        raise ValueError("work_areas must be [rows, cols, n_areas]")
    if inpainter_factory is None:
        if params is None:
            raise ValueError("either params or inpainter_factory is required")
        name = params.subparser_name.lower()
        if name not in FD_PDE_METHODS and name[0:4] != "ccst":
            raise ValueError("'%s' is not a finite-difference PDE inpainter" % params.subparser_name)
        inpainter_factory = lambda: create_inpainter_from_params(params)  # noqa: E731

    elevation_int = np.copy(elevation)                                   # interpolate_netcdf4.py:56
    for i in range(work_areas.shape[2]):
        cur_work_area = work_areas[:, :, i]
        rows = np.any(cur_work_area, axis=1)
        cols = np.any(cur_work_area, axis=0)
        if not rows.any():
            continue
        rmin, rmax = np.where(rows)[0][[0, -1]]
        cmin, cmax = np.where(cols)[0][[0, -1]]
        # True == known for the inpainters (inverse of mask_int); cells of the box outside the area stay fixed
        cur_inpaint_mask = np.copy(~mask_int[rmin:rmax + 1, cmin:cmax + 1])
        cur_elevation = np.copy(elevation[rmin:rmax + 1, cmin:cmax + 1])
        cur_inpaint_mask = np.logical_or(cur_inpaint_mask, ~cur_work_area[rmin:rmax + 1, cmin:cmax + 1])
        cur_elevation[np.isnan(cur_elevation)] = 0
        inpainter = inpainter_factory()
        cur_elevation_int = inpainter.inpaint(cur_elevation, cur_inpaint_mask)
        elevation_slice = elevation_int[rmin:rmax + 1, cmin:cmax + 1]
        elevation_slice[~cur_inpaint_mask] = cur_elevation_int[~cur_inpaint_mask]
    return elevation_int
