"""Method-name dispatch of the reference CLI, for the B200 back end.

Mirrors ``create_inpainter_from_params`` / ``get_common_fd_pde_inpainters_params_from_args`` of
heightmap_interpolation/apps/apps_common.py:57-74,172-215: the same ``params`` namespace (argparse
result of ``interpolate_netcdf4 <method> ...``) builds the same inpainter classes, so the reference's
``interpolate()`` (apps/interpolate_netcdf4.py:200-204) can call ``inpainter.inpaint(...)`` unchanged.
INTEGRATION.md shows the three-line branch a maintainer adds for ``--backend b200``.
"""
from ..inpainting import AMLEInpainter, CCSTInpainter, SobolevInpainter, TVInpainter

FD_PDE_METHODS = ("harmonic", "tv", "ccst", "amle")


def get_common_fd_pde_inpainters_params_from_args(params):
    """Common options of all FD-PDE inpainters (apps_common.py:57-74); missing attributes fall back
    to the CLI defaults of apps_common.py:31-55."""
    g = lambda name, default: getattr(params, name, default)  # noqa: E731
    return {"update_step_size": params.update_step_size,
            "term_criteria": g("term_criteria", "absolute_percent"),
            "term_check_iters": g("term_check_iters", 1000),
            "term_thres": params.term_thres,
            "max_iters": g("max_iters", 1000000),
            "relaxation": g("relaxation", 0),
            "backend": g("backend", "b200"),
            "ti_arch": g("ti_arch", "gpu"),
            "print_progress": g("verbose", False),
            "print_progress_iters": g("print_progress_iters", 1000),
            "mgs_levels": g("mgs_levels", 1),
            "mgs_min_res": g("mgs_min_res", 100),
            "init_with": g("init_with", "nearest"),
            "convolver": g("convolver", "opencv"),
            "debug_dir": g("debug_dir", "")}


def create_inpainter_from_params(params):
    """Same dispatch as apps_common.py:172-215 for the FD-PDE methods: ``harmonic`` -> SobolevInpainter,
    ``tv`` -> TVInpainter, names starting with ``ccst`` -> CCSTInpainter, ``amle`` -> AMLEInpainter."""
    name = params.subparser_name.lower()
    if name in ("navier-stokes", "telea", "shiftmap", "ebi"):
        raise ValueError("'%s' is not a finite-difference PDE inpainter; the B200 back end covers "
                         "harmonic, tv, ccst and amle" % name)
    options = get_common_fd_pde_inpainters_params_from_args(params)
    if name == "harmonic":
        return SobolevInpainter(**options)
    if name == "tv":
        options["epsilon"] = params.epsilon
        return TVInpainter(**options)
    if name[0:4] == "ccst":
        options["tension"] = params.tension
        return CCSTInpainter(**options)
    if name == "amle":
        options["convolve_in_1d"] = getattr(params, "convolve_in_1d", False)
        return AMLEInpainter(**options)
    raise ValueError("unknown inpainting method '%s'" % params.subparser_name)
