import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import mygrad_b200 as mg

        return mg.device_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenFile:
    """lazy view of one tests/golden/*.npz with its JSON manifest"""

    def __init__(self, name: str):
        self.npz = np.load(GOLDEN / name, allow_pickle=False)
        self.manifest = json.loads(str(self.npz["manifest"])) if "manifest" in self.npz.files else {}

    def case(self, name: str) -> dict:
        pre = name + "/"
        return {k[len(pre):]: self.npz[k] for k in self.npz.files if k.startswith(pre)}

    def names(self):
        return list(self.manifest)


@pytest.fixture(scope="session")
def conv_golden():
    return GoldenFile("conv_cases.npz")


@pytest.fixture(scope="session")
def pool_golden():
    return GoldenFile("pool_cases.npz")


@pytest.fixture(scope="session")
def layer_golden():
    return GoldenFile("layer_cases.npz")


@pytest.fixture(scope="session")
def graph_golden():
    return GoldenFile("graph_cases.npz")


def tol(dtype):
    """north-star tolerances: rtol 1e-10 (fp64) / 1e-5 (fp32); atol scales with the data"""
    return (1e-10, 1e-10) if np.dtype(dtype) == np.float64 else (1e-5, 1e-5)


def assert_close(actual, desired, dtype=None, scale=None, err_msg=""):
    """rtol from the north star; atol = rtol * scale where scale defaults to max|desired|
    (pure rtol is ill-posed for sums that cancel to ~0, SURVEY.md §7 'hard parts')."""
    actual, desired = np.asarray(actual), np.asarray(desired)
    assert actual.shape == desired.shape, (actual.shape, desired.shape, err_msg)
    dt = dtype if dtype is not None else desired.dtype
    rtol, _ = tol(dt)
    if scale is None:
        scale = float(np.max(np.abs(desired))) if desired.size else 1.0
    np.testing.assert_allclose(actual, desired, rtol=rtol, atol=rtol * max(scale, 1e-30), err_msg=err_msg)
