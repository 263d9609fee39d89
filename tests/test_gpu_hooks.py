"""The reference's host-side hooks on the GPU: step_fun(f, mask) (fd_pde_inpainter.py:406-408 and the four
subclasses) and Convolver.__call__(image, kernel, mask) (convolver.py:28-29), through the C ABI
(hmi_step_fun_host, hmi_convolve3x3_host) — against the fixtures written by the unmodified reference
(tests/golden/step_fun.npz), the CPU oracle and SciPy's filters."""
import numpy as np
import pytest
import scipy.ndimage
import scipy.signal

from helpers import DTYPES, METHOD_OPTS, range_err
from oracle.fdpde_oracle import OracleInpainter

import heightmap_interpolation_b200 as hmi
from heightmap_interpolation_b200.inpainting.convolver import Convolver

pytestmark = pytest.mark.gpu

CLASSES = {"harmonic": hmi.SobolevInpainter, "ccst": hmi.CCSTInpainter, "tv": hmi.TVInpainter,
           "amle": hmi.AMLEInpainter}


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("conv", ["opencv", "masked", "scipy-ndimage", "scipy-signal"])
@pytest.mark.parametrize("method", list(CLASSES))
def test_step_fun_matches_reference_fixture(golden_step, method, conv, dt):
    """Same call the reference's own loop makes: step_fun(f, work mask), full work mask."""
    f = golden_step["field"].astype(DTYPES[dt])
    ref = golden_step["step__%s__%s__%s" % (method, conv, dt)]
    inp = CLASSES[method](convolver=conv, **METHOD_OPTS[method])
    got = inp.step_fun(f, np.ones(f.shape, dtype=bool))
    assert got.shape == f.shape and got.dtype == f.dtype
    assert range_err(got, ref) <= (1e-12 if dt == "f64" else 2e-6), (method, conv, dt)
    assert np.array_equal(inp.step_fun(f), got)            # no mask = everywhere


@pytest.mark.parametrize("method", list(CLASSES))
def test_step_fun_with_dilated_work_mask_equals_oracle_at_unknown_cells(golden_step, method):
    """The mask the reference really passes is the 3x3 dilation of the unknown set (:293-296): the step
    at every unknown cell is the true stencil value; cells outside the work mask return their input."""
    f = golden_step["field"].astype(np.float64)
    rng = np.random.default_rng(11)
    unknown = rng.random(f.shape) < 0.15
    work = scipy.ndimage.binary_dilation(unknown, structure=np.ones((3, 3), dtype=bool))
    inp = CLASSES[method](convolver="masked", **METHOD_OPTS[method])
    got = inp.step_fun(f, work)
    ref = OracleInpainter(method, convolver="masked", **METHOD_OPTS[method]).step_fun(f, work.astype(np.uint8))
    assert np.abs(got[unknown] - ref[unknown]).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(got[~work], f[~work])


def test_amle_step_fun_1d_branch_uses_symmetric_borders(golden_step):
    f = golden_step["field"].astype(np.float64)
    got = hmi.AMLEInpainter(convolve_in_1d=True, update_step_size=0.01).step_fun(f, None)
    ref = OracleInpainter("amle", convolve_in_1d=True, update_step_size=0.01).step_fun(f, None)
    assert range_err(got, ref) <= 1e-12


KERNELS = [np.array([[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]]),          # differential.py:28-30
           np.array([[0, 0, 0], [0, -1, 1], [0, 0, 0]]),                            # forward horizontal :13-15
           np.array([[0, -1, 0], [0, 1, 0], [0, 0, 0]]),                            # backward vertical :19-21
           np.arange(9, dtype=np.float64).reshape(3, 3) - 2.5]                      # asymmetric in both axes


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_convolver_call_equals_the_library_filters_the_reference_calls(dt):
    rng = np.random.default_rng(5)
    img = rng.normal(size=(57, 83)).astype(DTYPES[dt])
    mask = rng.random(img.shape) < 0.4
    tol = 1e-13 if dt == "f64" else 1e-6
    for k in KERNELS:
        # opencv = correlation with BORDER_REFLECT_101 (cv2.filter2D); SciPy's 'mirror' is that border rule
        want = scipy.ndimage.correlate(img.astype(np.float64), k, mode="mirror")
        got = Convolver("opencv")(img, k, None)
        assert got.dtype == img.dtype and range_err(got, want) <= tol
        # masked = true convolution, reflect-101, only where the mask is set, the input elsewhere
        want = scipy.ndimage.convolve(img.astype(np.float64), k, mode="mirror")
        for name in ("masked", "masked-parallel"):
            got = Convolver(name)(img, k, mask)
            assert range_err(got[mask], want[mask]) <= tol
            assert np.array_equal(got[~mask], img[~mask])
        # both SciPy names = scipy.signal.convolve(mode='same'): zero padding (convolver.py:15-18)
        want = scipy.signal.convolve(img.astype(np.float64), k, mode="same")
        for name in ("scipy-signal", "scipy-ndimage"):
            assert range_err(Convolver(name)(img, k, mask), want) <= tol


def test_convolver_call_rejects_what_it_cannot_filter():
    c = Convolver("masked")
    img = np.zeros((8, 8))
    with pytest.raises(ValueError):
        c(img, np.ones((5, 5)), np.ones((8, 8), dtype=bool))
    with pytest.raises(ValueError):
        c(img, np.ones((3, 3)), None)
    with pytest.raises(TypeError):
        Convolver("opencv")(np.zeros((8, 8), dtype=np.int32), np.ones((3, 3)), None)
