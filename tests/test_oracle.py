"""Pins the oracle: NumPy restatement and the plain-C second opinion against the
reference's by-hand known answers and the golden vectors generated from the live
reference (tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from conftest import GoldenFile, assert_close
from oracle import conv_pool_oracle as oc


@pytest.fixture(scope="module")
def direct():
    import subprocess
    from pathlib import Path

    root = Path(oc.__file__).resolve().parent
    subprocess.run(["make", "-C", str(root)], check=True, capture_output=True)
    return oc.DirectC()


# ---- by-hand known answers pinned by the reference's own tests -------------------------

def test_conv_known_answers_reference_tests():
    # tests/nnet/layers/test_conv.py:179-217
    x = np.arange(1.0, 5.0).reshape(1, 1, 4)
    k = -np.arange(1.0, 3.0).reshape(1, 1, 2)
    np.testing.assert_allclose(oc.conv_fwd(x, k, 2), [[[-5.0, -11.0]]])
    x = np.arange(1.0, 13.0).reshape(1, 1, 3, 4)
    k = -np.arange(1.0, 5.0).reshape(1, 1, 2, 2)
    np.testing.assert_allclose(oc.conv_fwd(x, k, (1, 2)), [[[[-44.0, -64.0], [-84.0, -104.0]]]])


def test_conv_docstring_example():
    # src/mygrad/nnet/layers/conv.py:251-264 (square wave * ones(5), stride 1)
    x = np.zeros((1, 1, 16))
    x[..., 5:11] = 1
    k = np.ones((1, 1, 5))
    y = oc.conv_fwd(x, k, 1)
    np.testing.assert_allclose(y[0, 0], [0, 1, 2, 3, 4, 5, 5, 4, 3, 2, 1, 0])
    g = np.ones_like(y)
    np.testing.assert_allclose(oc.conv_dx(g, x.shape, k, 1)[0, 0],
                               [1, 2, 3, 4, 5, 5, 5, 5, 5, 5, 5, 5, 4, 3, 2, 1])
    np.testing.assert_allclose(oc.conv_dw(g, x, k.shape, 1)[0, 0], [6, 6, 6, 6, 6])
    assert oc.conv_fwd(np.zeros((10, 3, 32, 32)), np.zeros((5, 3, 2, 2)), 2).shape == (10, 5, 16, 16)


POOL_X_A = np.array(
    [[[17, 10, 15, 28, 25, 23], [44, 26, 18, 16, 39, 34], [5, 42, 36, 0, 2, 46], [30, 20, 1, 31, 35, 43]],
     [[6, 7, 45, 27, 11, 8], [37, 4, 41, 22, 9, 33], [47, 3, 13, 32, 21, 38], [19, 12, 40, 24, 14, 29]]],
    dtype=np.float32)
POOL_X_B = np.array(
    [[[33, 27, 47, 11, 7, 36], [20, 18, 9, 2, 3, 17], [45, 31, 24, 12, 25, 19], [28, 1, 8, 16, 34, 14]],
     [[37, 39, 40, 41, 21, 13], [35, 15, 6, 4, 23, 30], [43, 46, 32, 10, 26, 42], [38, 5, 44, 29, 0, 22]]],
    dtype=np.float32)

# (x, pool, stride, upstream grad or None for ones, forward answer, backward answer)
# values: tests/nnet/layers/test_maxpool.py:33-83, 86-131, 134-178
POOL_KNOWN = [
    (POOL_X_A, (3,), (1,), "arange",
     [[[17, 28, 28, 28], [44, 26, 39, 39], [42, 42, 36, 46], [30, 31, 35, 43]],
      [[45, 45, 45, 27], [41, 41, 41, 33], [47, 32, 32, 38], [40, 40, 40, 29]]],
     [[[0, 0, 0, 6, 0, 0], [4, 5, 0, 0, 13, 0], [0, 17, 10, 0, 0, 11], [12, 0, 0, 13, 14, 15]],
      [[0, 0, 51, 19, 0, 0], [0, 0, 63, 0, 0, 23], [24, 0, 0, 51, 0, 27], [0, 0, 87, 0, 0, 31]]]),
    (POOL_X_A, (2, 3), (2, 1), "ones",
     [[[44, 28, 39, 39], [42, 42, 36, 46]], [[45, 45, 45, 33], [47, 40, 40, 38]]],
     [[[0, 0, 0, 1, 0, 0], [1, 0, 0, 0, 2, 0], [0, 2, 1, 0, 0, 1], [0, 0, 0, 0, 0, 0]],
      [[0, 0, 3, 0, 0, 0], [0, 0, 0, 0, 0, 1], [1, 0, 0, 0, 0, 1], [0, 0, 2, 0, 0, 0]]]),
    (POOL_X_B, (2, 2, 2), (1, 1, 2), "arange",
     [[[39, 47, 36], [46, 32, 42], [46, 44, 42]]],
     [[[0, 0, 1, 0, 0, 2], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]],
      [[0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [0, 9, 4, 0, 0, 13], [0, 0, 7, 0, 0, 0]]]),
]


@pytest.mark.parametrize("case", range(len(POOL_KNOWN)))
def test_pool_known_answers_reference_tests(case, direct):
    x, pool, stride, gkind, fwd, bwd = POOL_KNOWN[case]
    y = oc.max_pool_fwd(x, pool, stride)
    np.testing.assert_array_equal(y, np.array(fwd, dtype=np.float32))
    g = np.arange(y.size, dtype=np.float32).reshape(y.shape) if gkind == "arange" else np.ones_like(y)
    np.testing.assert_array_equal(oc.max_pool_bwd(g, x, pool, stride), np.array(bwd, dtype=np.float64))
    np.testing.assert_array_equal(direct.max_pool_fwd(x, pool, stride), np.array(fwd, dtype=np.float64))
    np.testing.assert_array_equal(direct.max_pool_bwd(g, x, pool, stride), np.array(bwd, dtype=np.float64))


def test_pool_docstring_example():
    # src/mygrad/nnet/layers/pooling.py:203-214
    x = np.array([[0.0, 10.0, 8.0], [9.0, 7.0, 3.0], [5.0, 0.0, 20.0]])
    y = oc.max_pool_fwd(x, (2,), 1)
    np.testing.assert_array_equal(y, [[10, 10], [9, 7], [5, 20]])
    np.testing.assert_array_equal(oc.max_pool_bwd(np.ones_like(y), x, (2,), 1),
                                  [[0, 2, 0], [1, 1, 0], [1, 0, 1]])
    assert oc.max_pool_fwd(np.zeros((10, 3, 12, 12)), (2, 2), (2, 1)).shape == (10, 3, 6, 11)


# ---- golden vectors generated from the live reference ------------------------------------

def _conv_names():
    from conftest import GoldenFile
    return GoldenFile("conv_cases.npz").names()


@pytest.mark.parametrize("name", _conv_names())
def test_conv_oracle_vs_golden(name, conv_golden, direct):
    c, m = conv_golden.case(name), conv_golden.manifest[name]
    s, p, d = m["stride"], m["padding"], m["dilation"]
    x, w, g = c["x"], c["w"], c["g"]
    ydt = np.result_type(x.dtype, w.dtype)
    y = oc.conv_fwd(x, w, s, p, d)
    assert y.dtype == c["y"].dtype and y.shape == c["y"].shape
    assert_close(y, c["y"], ydt, err_msg=name)
    dx = oc.conv_dx(g, x.shape, w, s, p, d, dtype=x.dtype)
    assert dx.dtype == c["dx"].dtype
    assert_close(dx, c["dx"], np.result_type(ydt, x.dtype) if x.dtype == np.float64 else np.float32, err_msg=name)
    dw = oc.conv_dw(g, x, w.shape, s, p, d).astype(w.dtype)
    assert_close(dw, c["dw"], ydt if w.dtype == np.float64 else np.float32, err_msg=name)
    # plain-C second opinion (float64 accumulation)
    assert_close(direct.conv_fwd(x, w, s, p, d), c["y"], ydt, err_msg=name)
    assert_close(direct.conv_dx(g, x.shape, w, s, p, d), c["dx"], c["dx"].dtype, err_msg=name)
    assert_close(direct.conv_dw(g, x, w.shape, s, p, d), c["dw"], c["dw"].dtype, err_msg=name)


def _pool_names():
    from conftest import GoldenFile
    return GoldenFile("pool_cases.npz").names()


@pytest.mark.parametrize("name", _pool_names())
def test_pool_oracle_vs_golden_bit_exact(name, pool_golden, direct):
    c, m = pool_golden.case(name), pool_golden.manifest[name]
    x, g = c["x"], c["g"]
    y = oc.max_pool_fwd(x, m["pool"], m["stride"])
    np.testing.assert_array_equal(y, c["y"])           # NaN == NaN under assert_array_equal
    dx = oc.max_pool_bwd(g, x, m["pool"], m["stride"]).astype(x.dtype)
    np.testing.assert_array_equal(dx, c["dx"])
    np.testing.assert_array_equal(direct.max_pool_fwd(x, m["pool"], m["stride"]).astype(x.dtype), c["y"])
    np.testing.assert_array_equal(direct.max_pool_bwd(g, x, m["pool"], m["stride"]).astype(x.dtype), c["dx"])
    rtol = 1e-10 if x.dtype == np.float64 else 1e-5
    with np.errstate(invalid="ignore"):
        np.testing.assert_allclose(oc.mean_pool_fwd(x, m["pool"], m["stride"]), c["mean"],
                                   rtol=rtol, atol=rtol, equal_nan=True)
    if "mean_dx" in c:    # the oracle's mean-pool gradient against the reference's own autodiff (a7)
        np.testing.assert_allclose(oc.mean_pool_bwd(g, x.shape, m["pool"], m["stride"]).astype(x.dtype),
                                   c["mean_dx"], rtol=rtol, atol=rtol)


def _layer_names():
    from conftest import GoldenFile
    return GoldenFile("layer_cases.npz").names()


@pytest.mark.parametrize("name", _layer_names())
def test_layer_step_vs_golden(name, layer_golden):
    c, m = layer_golden.case(name), layer_golden.manifest[name]
    out, dx, dw = oc.layer_step(c["x"], c["w"], stride=m["stride"], padding=m["padding"],
                                dilation=m["dilation"], pool=m["pool"], pool_stride=m["pool_stride"])
    dt = c["x"].dtype
    assert_close(out, c["out"], dt, err_msg=name)
    assert_close(dx, c["dx"], dt, err_msg=name)
    assert_close(dw, c["dw"], dt, err_msg=name)


def test_reduce_broadcast_vs_golden(graph_golden):
    n = int(graph_golden.npz["rb_count"])
    for i in range(n):
        c = graph_golden.case(f"rb{i}")
        out = oc.reduce_broadcast(c["g"], tuple(int(v) for v in c["vshape"]))
        assert out.shape == c["out"].shape
        assert_close(out, c["out"], np.float64)


def test_oracle_rejects_incompatible_shapes():
    with pytest.raises(ValueError):
        oc.conv_fwd(np.zeros((1, 2, 64, 64)), np.zeros((3, 2, 5, 5)), 2, 0, 2)   # BASELINE C3 as written
    assert oc.conv_fwd(np.zeros((1, 2, 65, 65)), np.zeros((3, 2, 5, 5)), 2, 0, 2).shape == (1, 3, 29, 29)
    with pytest.raises(ValueError):
        oc.max_pool_fwd(np.zeros((1, 2, 2, 2)), (3, 3, 3), 1)


# ---- batchnorm -------------------------------------------------------------

def test_batchnorm_docstring_example():
    # src/mygrad/nnet/layers/batchnorm.py:152-157
    x = np.array([1.0, 4.0, 1.0]).reshape(3, 1)
    y, mean, var = oc.batchnorm_fwd(x, None, None, 0.0)
    np.testing.assert_allclose(y[:, 0], [-0.70710678, 1.41421356, -0.70710678], rtol=1e-8)
    np.testing.assert_allclose(mean, [2.0])
    np.testing.assert_allclose(var, [2.0])


@pytest.mark.parametrize("name", GoldenFile("batchnorm_cases.npz").names())
def test_batchnorm_oracle_vs_golden(name):
    gf = GoldenFile("batchnorm_cases.npz")
    c, eps = gf.case(name), gf.manifest[name]["eps"]
    x, g = c["x"], c["g"]
    dt = x.dtype
    y, mean, var = oc.batchnorm_fwd(x, c.get("gamma"), c.get("beta"), eps)
    # the offset-mean case amplifies rounding of the mean by mean/std = 1e6
    loose = 1e6 if name == "bn_offset_mean" else 1.0
    tol = dict(rtol=(1e-10 if dt == np.float64 else 2e-5) * loose, atol=(1e-12 if dt == np.float64 else 1e-6) * loose)
    np.testing.assert_allclose(y, c["y"], **tol)
    np.testing.assert_allclose(mean, c["mean"], rtol=1e-12 if dt == np.float64 else 1e-6)
    np.testing.assert_allclose(var, c["var"], rtol=1e-9 if dt == np.float64 else 1e-5)
    dx, dgamma, dbeta = oc.batchnorm_bwd(g, x, c.get("gamma"), eps)
    scale = float(np.abs(c["dx"]).max()) or 1.0
    np.testing.assert_allclose(dx, c["dx"], rtol=tol["rtol"], atol=tol["rtol"] * scale * 10)
    if "dgamma" in c:
        np.testing.assert_allclose(dgamma, c["dgamma"], rtol=tol["rtol"] * 10, atol=tol["atol"] * 100)
    if "dbeta" in c:
        np.testing.assert_allclose(dbeta, c["dbeta"], rtol=tol["rtol"] * 10, atol=tol["atol"] * 100)
