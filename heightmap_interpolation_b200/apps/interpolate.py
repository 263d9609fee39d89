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

Two routes.  When the inpainter's initialiser exists on the device (``zeros``, ``mean``; single level) the
grid is uploaded once and everything per area — crop, known mask, NaN -> 0, initialiser, solve, paste of
the inpainted cells — runs on the GPU(s) (csrc/hmi_areas.cu); elevation_int comes back once at the end.
Otherwise (``nearest`` ... ``sobolev`` initialisers, multigrid, foreign inpainter objects) each area goes
through ``inpainter.inpaint(window, mask)`` on host arrays, as in the reference.
"""
import numpy as np

from .apps_common import FD_PDE_METHODS, create_inpainter_from_params


def _bounding_box(layer):
    """First / last row and column holding a True (interpolate_netcdf4.py:181-185), or None if empty."""
    rows = np.flatnonzero(layer.any(axis=1))
    if rows.size == 0:
        return None
    cols = np.flatnonzero(layer.any(axis=0))
    return int(rows[0]), int(rows[-1]) + 1, int(cols[0]), int(cols[-1]) + 1


def inpaint_work_areas(elevation, mask_int, work_areas=None, params=None, inpainter_factory=None, on_device=None):
    """Returns ``elevation_int``: a copy of ``elevation`` with the cells of ``mask_int`` (True == to be
    interpolated) that fall inside each work area replaced by the inpainted values.

    ``work_areas``: bool array [rows, cols, n_areas] (None = one area covering the grid, what the
    reference builds without ``--areas``, netcdf_data_io.py:262).  ``inpainter_factory`` overrides
    ``create_inpainter_from_params(params)`` (a fresh inpainter per area, like the reference).
    ``on_device``: None = use the device route whenever the inpainter supports it, False = always go
    through host arrays, True = require the device route.
    """
    elevation = np.asarray(elevation)
    mask_int = np.asarray(mask_int, dtype=bool)
    if elevation.ndim != 2 or mask_int.shape != elevation.shape:
        raise ValueError("elevation and mask_int must be 2-D arrays of the same shape")
    if work_areas is not None:
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

    layers = [None] if work_areas is None else [work_areas[:, :, i] for i in range(work_areas.shape[2])]
    probe = inpainter_factory()
    device_ok = (elevation.dtype in (np.float32, np.float64) and hasattr(probe, "inpaint_window")
                 and probe.can_inpaint_window())
    if on_device and not device_ok:
        raise ValueError("this inpainter / grid cannot take the device route (initialiser, multigrid or dtype)")
    if device_ok and on_device is not False:
        return _areas_on_device(elevation, mask_int, layers, inpainter_factory, probe)
    return _areas_through_host(elevation, mask_int, layers, inpainter_factory)


def _areas_on_device(elevation, mask_int, layers, inpainter_factory, first):
    from ..solver import AreaSession
    with AreaSession(elevation, mask_int, device=first._device()) as session:
        for i, layer in enumerate(layers):
            box = (0, elevation.shape[0], 0, elevation.shape[1]) if layer is None else _bounding_box(layer)
            if box is None:
                continue
            session.set_work_area(layer)
            inpainter = first if i == 0 else inpainter_factory()
            inpainter.inpaint_window(session, box[0], box[2], box[1] - box[0], box[3] - box[2])
        return session.result()


def _areas_through_host(elevation, mask_int, layers, inpainter_factory):
    elevation_int = elevation.copy()
    for layer in layers:
        box = (0, elevation.shape[0], 0, elevation.shape[1]) if layer is None else _bounding_box(layer)
        if box is None:
            continue
        window = (slice(box[0], box[1]), slice(box[2], box[3]))
        known = ~mask_int[window]                       # True == known for the inpainters
        if layer is not None:
            known |= ~layer[window]                     # cells of the box outside the area stay fixed
        grid = np.nan_to_num(elevation[window], copy=True, nan=0.0, posinf=np.inf, neginf=-np.inf)
        result = inpainter_factory().inpaint(grid, known)
        np.copyto(elevation_int[window], result, where=~known)
    return elevation_int
