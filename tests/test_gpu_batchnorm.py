"""Parity of the device batchnorm with the reference: golden vectors from the live
reference (tests/golden/make_golden.py), seeded random cases against the oracle, and the drop-in
surface of ``mygrad.nnet.layers.batchnorm`` (src/mygrad/nnet/layers/batchnorm.py:13-175).
Tolerance: rtol 1e-10 (f64) / 1e-5 (f32) on max|value|, as for the convolutions."""
import numpy as np
import pytest

import mygrad_b200 as mg
from conftest import GoldenFile, assert_close
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import BatchNorm, batchnorm, conv_nd, max_pool
from oracle import conv_pool_oracle as oc

pytestmark = pytest.mark.gpu


def test_batchnorm_docstring_example():
    # batchnorm.py:152-157
    x = Tensor(np.array([1.0, 4.0, 1.0]).reshape(3, 1))
    y = batchnorm(x, eps=0)
    np.testing.assert_allclose(y.numpy()[:, 0], [-0.70710678, 1.41421356, -0.70710678], rtol=1e-8)
    assert isinstance(y.creator, BatchNorm)
    np.testing.assert_allclose(y.creator.mean.get(), [2.0])
    np.testing.assert_allclose(y.creator.var.get(), [2.0])


@pytest.mark.parametrize("name", GoldenFile("batchnorm_cases.npz").names())
def test_batchnorm_golden(name):
    gf = GoldenFile("batchnorm_cases.npz")
    c, eps = gf.case(name), gf.manifest[name]["eps"]
    x, g = c["x"], c["g"]
    dt = x.dtype
    xt = Tensor(x)
    gamma = Tensor(c["gamma"]) if "gamma" in c else None
    beta = Tensor(c["beta"]) if "beta" in c else None
    y = batchnorm(xt, gamma=gamma, beta=beta, eps=eps)
    assert y.shape == x.shape and y.dtype == dt
    # with mean/std = 1e6 (bn_offset_mean) one rounding of the mean is already 1e6 * eps of the output
    loose = 1e6 if name == "bn_offset_mean" else 1.0
    scale = float(np.abs(c["y"]).max()) * loose
    assert_close(y.numpy(), c["y"], dt, scale=scale, err_msg=f"{name} y")
    # a mean is only defined to eps * (|mean| + std): the reference's own float32 pairwise sum is that far off
    spread = float(np.abs(c["mean"]).max() + np.sqrt(c["var"].max()))
    assert_close(y.creator.mean.get(), c["mean"], dt, scale=spread, err_msg=f"{name} mean")
    assert_close(y.creator.var.get(), c["var"], dt, err_msg=f"{name} var")
    y.backward(g)
    assert xt.grad.dtype == dt
    assert_close(xt.grad.get(), c["dx"], dt, scale=float(np.abs(c["dx"]).max()) * loose * 10, err_msg=f"{name} dx")
    if gamma is not None:
        assert gamma.grad.shape == c["dgamma"].shape
        assert_close(gamma.grad.get(), c["dgamma"], dt, scale=float(np.abs(g).sum() / g.shape[1]) * loose,
                     err_msg=f"{name} dgamma")
    if beta is not None:
        assert_close(beta.grad.get(), c["dbeta"], dt, scale=float(np.abs(g).sum() / g.shape[1]),
                     err_msg=f"{name} dbeta")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_batchnorm_random_vs_oracle(dtype):
    rng = np.random.default_rng(5)
    for _ in range(12):
        nd = int(rng.integers(2, 6))
        shape = tuple(int(rng.integers(1, 7)) for _ in range(nd))
        if np.prod(shape) // shape[1] < 2:
            continue
        x = (rng.standard_normal(shape) * rng.uniform(0.1, 3) + rng.uniform(-2, 2)).astype(dtype)
        gamma = rng.standard_normal(shape[1]).astype(dtype)
        beta = rng.standard_normal(shape[1]).astype(dtype)
        eps = float(rng.choice([1e-8, 1e-5, 1e-2]))
        xt, gt, bt = Tensor(x), Tensor(gamma), Tensor(beta)
        y = batchnorm(xt, gamma=gt, beta=bt, eps=eps)
        want, _, _ = oc.batchnorm_fwd(x.astype(np.float64), gamma, beta, eps)
        assert_close(y.numpy(), want, dtype, scale=float(np.abs(want).max()) * 10)
        g = rng.standard_normal(shape).astype(dtype)
        y.backward(g)
        dx, dgamma, dbeta = oc.batchnorm_bwd(g.astype(np.float64), x.astype(np.float64), gamma, eps)
        n = x.size / shape[1]
        assert_close(xt.grad.get(), dx, dtype, scale=float(np.abs(dx).max()) * 50 + 1e-3)
        assert_close(gt.grad.get(), dgamma, dtype, scale=float(np.abs(g).max()) * n * 10)
        assert_close(bt.grad.get(), dbeta, dtype, scale=float(np.abs(g).max()) * n)


def test_batchnorm_constant_and_errors():
    x = Tensor(np.random.default_rng(0).standard_normal((4, 3, 5)))
    assert batchnorm(x, eps=1e-5, constant=True).constant is True
    with pytest.raises(TypeError):
        batchnorm(x)                                  # eps is required (batchnorm.py:119)
    with pytest.raises(ValueError):
        batchnorm(x, gamma=np.ones(4), eps=1e-5)      # gamma must be (C,)
    # gamma / beta passed as plain arrays are constants: no gradient for them, x still gets one
    y = batchnorm(x, gamma=np.ones(3), beta=np.zeros(3), eps=1e-5)
    y.backward()
    assert x.grad is not None and x.grad.shape == x.shape
    # sum_n of a batch-normalised tensor has zero gradient w.r.t. x (mean removal)
    np.testing.assert_allclose(x.grad.get(), 0.0, atol=1e-12)


def test_batchnorm_in_a_layer_stack():
    """conv -> batchnorm -> max_pool through Tensor.backward against the oracle chain (f32)"""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((6, 4, 12, 12)).astype(np.float32)
    w = rng.standard_normal((8, 4, 3, 3)).astype(np.float32)
    gamma = (rng.standard_normal(8) + 1).astype(np.float32)
    beta = rng.standard_normal(8).astype(np.float32)
    xt, wt, gt, bt = Tensor(x), Tensor(w), Tensor(gamma), Tensor(beta)
    y = conv_nd(xt, wt, stride=1, padding=1)
    z = batchnorm(y, gamma=gt, beta=bt, eps=1e-5)
    out = max_pool(z, (2, 2), 2)
    out.backward()
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    y_ref = oc.conv_fwd(x64, w64, 1, 1, 1)
    z_ref, _, _ = oc.batchnorm_fwd(y_ref, gamma, beta, 1e-5)
    assert_close(z.numpy(), z_ref, np.float32, scale=float(np.abs(z_ref).max()) * 4)
    # route the pooled ones through the GPU's own z (pool routing is exact on its input)
    dz = oc.max_pool_bwd(np.ones(out.shape), z.numpy().astype(np.float64), (2, 2), 2)
    dy, dgamma, dbeta = oc.batchnorm_bwd(dz, y_ref, gamma, 1e-5)
    assert_close(gt.grad.get(), dgamma, np.float32, scale=float(np.abs(dz).sum() / 8) * 4)
    assert_close(bt.grad.get(), dbeta, np.float32, scale=float(np.abs(dz).sum() / 8))
    assert_close(wt.grad.get(), oc.conv_dw(dy, x64, w.shape, 1, 1, 1), np.float32,
                 scale=float(np.abs(dy).sum() / 8))
    assert_close(xt.grad.get(), oc.conv_dx(dy, x.shape, w64, 1, 1, 1), np.float32,
                 scale=float(np.abs(dy).max()) * 100)


def test_batchnorm_full_size_properties():
    """C2-sized activation (256, 128, 56, 56) f32: per-channel mean 0 / variance 1 of the output,
    beta gradient = per-channel sum of the upstream gradient, dx orthogonal to 1 and to x_norm"""
    rng = np.random.default_rng(1)
    shape = (64, 128, 56, 56)
    x = mg.asdevice((rng.standard_normal(shape) * 2 + 0.5).astype(np.float32))
    gamma = Tensor((rng.standard_normal(128) + 1.5).astype(np.float32))
    beta = Tensor(rng.standard_normal(128).astype(np.float32))
    xt = Tensor(x, copy=False)
    y = batchnorm(xt, gamma=gamma, beta=beta, eps=1e-5)
    yh = y.numpy().astype(np.float64)
    g_, b_ = gamma.numpy().astype(np.float64), beta.numpy().astype(np.float64)
    xn = (yh - b_[None, :, None, None]) / g_[None, :, None, None]
    np.testing.assert_allclose(xn.mean(axis=(0, 2, 3)), 0.0, atol=2e-5)
    np.testing.assert_allclose(xn.var(axis=(0, 2, 3)), 1.0, rtol=1e-4)
    g = rng.standard_normal(shape).astype(np.float32)
    y.backward(g)
    g64 = g.astype(np.float64)
    np.testing.assert_allclose(beta.grad.get(), g64.sum(axis=(0, 2, 3)), rtol=1e-5, atol=1e-2)
    np.testing.assert_allclose(gamma.grad.get(), (g64 * xn).sum(axis=(0, 2, 3)), rtol=1e-4, atol=0.5)
    dx = xt.grad.get().astype(np.float64)
    n = dx.size / 128
    assert np.all(np.abs(dx.sum(axis=(0, 2, 3))) / n < 1e-6)
    assert np.all(np.abs((dx * xn).sum(axis=(0, 2, 3))) / n < 1e-5)
