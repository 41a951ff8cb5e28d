"""save / load in the reference's .npz format and the public sliding_window_view on device arrays
(SURVEY.md §8f-4): fixtures written by the live reference (tests/golden/make_golden.py)."""
import io

import numpy as np
import pytest

import mygrad_b200 as mg
from conftest import GOLDEN, GoldenFile
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers.utils import sliding_window_view

pytestmark = pytest.mark.gpu


def test_load_file_written_by_the_reference():
    # src/mygrad/_io.py:41-61 docstring example: x = arange(10.), (x * x).backward()
    t = mg.load(GOLDEN / "saved_by_reference.npz")
    assert isinstance(t, Tensor) and t.dtype == np.float64
    np.testing.assert_array_equal(t.numpy(), np.arange(10.0))
    np.testing.assert_array_equal(t.grad.get(), 2 * np.arange(10.0))
    t = mg.load(GOLDEN / "saved_by_reference_nograd.npz")
    assert t.grad is None and t.dtype == np.float32 and t.shape == (3, 4)


def test_save_round_trip_and_keys():
    rng = np.random.default_rng(3)
    x = Tensor(rng.standard_normal((5, 7)).astype(np.float32))
    buf = io.BytesIO()
    mg.save(buf, x)
    buf.seek(0)
    assert sorted(np.load(buf).files) == ["data"]              # no grad key without a gradient (_io.py:60-63)
    g = rng.standard_normal((5, 7)).astype(np.float32)
    x.backward(g)
    buf = io.BytesIO()
    mg.save(buf, x)
    buf.seek(0)
    raw = np.load(buf)
    assert sorted(raw.files) == ["data", "grad"]
    np.testing.assert_array_equal(raw["data"], x.numpy())
    np.testing.assert_array_equal(raw["grad"], g)
    buf.seek(0)
    y = mg.load(buf)
    np.testing.assert_array_equal(y.numpy(), x.numpy())
    np.testing.assert_array_equal(y.grad.get(), g)
    with pytest.raises(TypeError):
        mg.save(io.BytesIO(), x.numpy())                          # _io.py:55-58


@pytest.mark.parametrize("name", GoldenFile("window_cases.npz").names())
def test_sliding_window_view_golden(name):
    gf = GoldenFile("window_cases.npz")
    c, m = gf.case(name), gf.manifest[name]
    step = m["step"] if isinstance(m["step"], int) else tuple(m["step"])
    dil = m["dilation"] if (m["dilation"] is None or isinstance(m["dilation"], int)) else tuple(m["dilation"])
    view = sliding_window_view(mg.asdevice(c["x"]), tuple(m["window_shape"]), step, dil)
    assert view.shape == c["view"].shape and view.dtype == c["x"].dtype
    np.testing.assert_array_equal(view.get(), c["view"])           # a gather: bit-exact


def test_sliding_window_view_errors():
    x = mg.asdevice(np.zeros((4, 5)))
    with pytest.raises(TypeError):
        sliding_window_view(x, 2, 1)                                # utils.py:123-126
    with pytest.raises(TypeError):
        sliding_window_view(x, (2.0,), 1)                           # utils.py:128-132
    with pytest.raises(ValueError):
        sliding_window_view(x, (1, 1, 1), 1)                        # utils.py:134-138
    with pytest.raises(ValueError):
        sliding_window_view(x, (2, 2), 0)                           # utils.py:150-154
    with pytest.raises(ValueError):
        sliding_window_view(x, (5, 6), 1)                           # utils.py:156-161
    with pytest.raises(ValueError):
        sliding_window_view(x, (2, 2), 1, dilation=(1, 3))          # utils.py:186-193
    with pytest.raises(TypeError):
        sliding_window_view(x, (2, 2), 1, dilation=1.5)             # utils.py:163-171
