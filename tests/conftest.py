import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_fk():
    return dict(np.load(os.path.join(GOLDEN, "fk.npz")))


@pytest.fixture(scope="session")
def golden_dq():
    return dict(np.load(os.path.join(GOLDEN, "dq.npz")))


@pytest.fixture(scope="session")
def golden_quat():
    return dict(np.load(os.path.join(GOLDEN, "quat.npz")))


@pytest.fixture(scope="session")
def golden_quat_ext():
    return dict(np.load(os.path.join(GOLDEN, "quat_ext.npz")))


@pytest.fixture(scope="session")
def golden_ik():
    return dict(np.load(os.path.join(GOLDEN, "ik.npz")))


@pytest.fixture(scope="session")
def golden_misc():
    return dict(np.load(os.path.join(GOLDEN, "misc.npz")))


@pytest.fixture(scope="session")
def golden_bvh():
    return dict(np.load(os.path.join(GOLDEN, "bvh.npz")))


@pytest.fixture
def set_knobs(monkeypatch):
    """Force kernel variants for one test: the library honours its PMB_* experiment knobs only under
    PMB_EXPERIMENT=1 and reads them once, so set the environment and ask it to re-read (pmb_reload_knobs)."""
    from pymotion_b200 import _lib

    def apply(knobs: dict):
        monkeypatch.setenv("PMB_EXPERIMENT", "1")
        for k, v in knobs.items():
            monkeypatch.setenv(k, str(v))
        _lib.load().pmb_reload_knobs()

    yield apply
    monkeypatch.undo()
    _lib.load().pmb_reload_knobs()
