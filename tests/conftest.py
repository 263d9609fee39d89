import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_step():
    return np.load(os.path.join(GOLDEN, "step_fun.npz"))


@pytest.fixture(scope="session")
def golden_sweeps():
    return np.load(os.path.join(GOLDEN, "sweeps.npz"))


@pytest.fixture(scope="session")
def golden_converged():
    return np.load(os.path.join(GOLDEN, "converged.npz"))


@pytest.fixture(scope="session")
def golden_multigrid():
    return np.load(os.path.join(GOLDEN, "multigrid.npz"))


@pytest.fixture(scope="session")
def golden_nearest():
    return np.load(os.path.join(GOLDEN, "nearest.npz"))
