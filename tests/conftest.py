import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_spielberg():
    return np.load(os.path.join(GOLDEN, "pp_spielberg.npz"))


@pytest.fixture(scope="session")
def golden_ellipse():
    return np.load(os.path.join(GOLDEN, "pp_ellipse.npz"))


@pytest.fixture(scope="session")
def golden_misc():
    return np.load(os.path.join(GOLDEN, "misc.npz"))


@pytest.fixture(scope="session")
def ellipse():
    from f1tenth_planning_b200 import synth
    return synth.ellipse_track()


@pytest.fixture(scope="session")
def corridor():
    from f1tenth_planning_b200 import synth
    return synth.corridor_grid()
