"""Parity of the CUDA path with the reference, through the public API
(conv_nd / max_pool / Tensor.backward -> ctypes -> libmgb200.so).

Checks, in order: the reference's by-hand known answers, the golden vectors generated
from the live reference, seeded random cases against the oracle, and — at BASELINE.json's
full sizes — size-independent properties plus spot checks of single samples.
Tolerances are the north star's: pooling bit-exact, conv rtol 1e-10 (f64) / 1e-5 (f32).
"""
import numpy as np
import pytest

import mygrad_b200 as mg
from conftest import GoldenFile, assert_close
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, max_pool, mean_pool
from oracle import conv_pool_oracle as oc
from test_oracle import POOL_KNOWN

pytestmark = pytest.mark.gpu


# ---- known answers pinned by the reference's own tests -------------------------------

def test_conv_known_answers():
    # tests/nnet/layers/test_conv.py:179-217
    x = Tensor(np.arange(1, 5).reshape(1, 1, 4).astype(float))
    k = Tensor(-1 * np.arange(1, 3).reshape(1, 1, 2).astype(float))
    o = conv_nd(x, k, stride=(2,), constant=True)
    assert isinstance(o, Tensor) and o.constant is True
    np.testing.assert_allclose(o.numpy(), [[[-5.0, -11.0]]])
    x = Tensor(np.arange(1, 13).reshape(1, 1, 3, 4).astype(float))
    k = Tensor(-1 * np.arange(1, 5).reshape(1, 1, 2, 2).astype(float))
    o = conv_nd(Tensor(x), k, stride=(1, 2), constant=True)
    np.testing.assert_allclose(o.numpy(), [[[[-44.0, -64.0], [-84.0, -104.0]]]])


def test_conv_docstring_example():
    # src/mygrad/nnet/layers/conv.py:251-264
    xh = np.zeros((1, 1, 16), dtype=np.float32)
    xh[..., 5:11] = 1
    x, k = Tensor(xh), Tensor(np.ones((1, 1, 5), dtype=np.float32))
    y = conv_nd(x, k, stride=1)
    assert y.dtype == np.float32
    np.testing.assert_allclose(y.numpy()[0, 0], [0, 1, 2, 3, 4, 5, 5, 4, 3, 2, 1, 0])
    y.backward()
    np.testing.assert_allclose(x.grad.get()[0, 0], [1, 2, 3, 4, 5, 5, 5, 5, 5, 5, 5, 5, 4, 3, 2, 1])
    np.testing.assert_allclose(k.grad.get()[0, 0], [6, 6, 6, 6, 6])
    assert conv_nd(np.zeros((10, 3, 32, 32)), np.zeros((5, 3, 2, 2)), stride=2).shape == (10, 5, 16, 16)


@pytest.mark.parametrize("case", range(len(POOL_KNOWN)))
def test_pool_known_answers(case):
    # tests/nnet/layers/test_maxpool.py:33-178
    xh, pool, stride, gkind, fwd, bwd = POOL_KNOWN[case]
    x = Tensor(xh, dtype="float32")
    out = max_pool(x, pool, stride)
    g = np.arange(out.size).reshape(out.shape) if gkind == "arange" else None
    out.backward(g)
    np.testing.assert_array_equal(out.numpy(), np.array(fwd, dtype=np.float32))
    assert x.grad.dtype == np.float32
    np.testing.assert_array_equal(x.grad.get(), np.array(bwd, dtype=np.float32))
    assert max_pool(x, pool, stride, constant=True).constant is True
    assert max_pool(x, pool, stride, constant=False).constant is False


def test_pool_docstring_example():
    # src/mygrad/nnet/layers/pooling.py:203-214
    x = Tensor([[0.0, 10.0, 8.0], [9.0, 7.0, 3.0], [5.0, 0.0, 20.0]])
    out = max_pool(x, pool=(2,), stride=1)
    np.testing.assert_array_equal(out.numpy(), [[10, 10], [9, 7], [5, 20]])
    out.backward()
    np.testing.assert_array_equal(x.grad.get(), [[0, 2, 0], [1, 1, 0], [1, 0, 1]])
    assert max_pool(np.random.rand(10, 3, 12, 12), (2, 2), (2, 1)).shape == (10, 3, 6, 11)


def test_padding_only_input_gives_exact_zero():
    # tests/nnet/layers/test_conv.py:247-273
    for nd, padding in ((1, 2), (2, (1, 3)), (3, 1)):
        pads = (padding,) * nd if isinstance(padding, int) else padding
        x = Tensor(np.zeros((1, 1) + (0,) * nd))
        k = np.random.default_rng(0).standard_normal((1, 1) + tuple(2 * p for p in pads))
        out = conv_nd(x, k, padding=padding, stride=1)
        assert out.shape == (1,) * x.ndim
        assert out.item() == 0.0
        out.backward()
        assert x.grad.shape == x.shape


# ---- golden vectors generated from the live reference ---------------------------------

@pytest.mark.parametrize("name", GoldenFile("conv_cases.npz").names())
def test_conv_vs_golden(name, conv_golden):
    c, m = conv_golden.case(name), conv_golden.manifest[name]
    x, w = Tensor(c["x"]), Tensor(c["w"])
    y = conv_nd(x, w, stride=m["stride"], padding=m["padding"], dilation=m["dilation"])
    assert y.shape == c["y"].shape and y.dtype == c["y"].dtype
    y.backward(c["g"])
    assert_close(y.numpy(), c["y"], c["y"].dtype, err_msg=name)
    assert x.grad.dtype == c["dx"].dtype and w.grad.dtype == c["dw"].dtype
    assert_close(x.grad.get(), c["dx"], c["dx"].dtype, err_msg=name + " dx")
    assert_close(w.grad.get(), c["dw"], c["dw"].dtype, err_msg=name + " dw")
    assert y.creator is None and not x._ops      # graph cleared after backward


@pytest.mark.parametrize("name", GoldenFile("pool_cases.npz").names())
def test_pool_vs_golden_bit_exact(name, pool_golden):
    c, m = pool_golden.case(name), pool_golden.manifest[name]
    x = Tensor(c["x"])
    y = max_pool(x, tuple(m["pool"]), m["stride"])
    y.backward(c["g"])
    np.testing.assert_array_equal(y.numpy(), c["y"], err_msg=name)
    assert x.grad.dtype == c["x"].dtype
    np.testing.assert_array_equal(x.grad.get(), c["dx"], err_msg=name + " dx")
    # mean-pool: the reference has no such function.  Forward is pinned to the reference's
    # sliding_window_view(...).mean(), backward to the gradient the reference's OWN autodiff gives for
    # mean-pooling composed of its ops (strided Tensor slices, add, divide: make_golden.ref_mean_pool)
    xm = Tensor(c["x"])
    ym = mean_pool(xm, tuple(m["pool"]), m["stride"])
    rtol = 1e-10 if c["x"].dtype == np.float64 else 1e-5
    np.testing.assert_allclose(ym.numpy(), c["mean"], rtol=rtol, atol=rtol, equal_nan=True, err_msg=name)
    if np.isfinite(c["x"]).all():
        ym.backward(c["g"])
        assert xm.grad.dtype == c["x"].dtype
        np.testing.assert_allclose(xm.grad.get(), c["mean_dx"], rtol=rtol, atol=rtol, err_msg=name + " mean dx")


@pytest.mark.parametrize("name", GoldenFile("layer_cases.npz").names())
def test_layer_vs_golden(name, layer_golden):
    c, m = layer_golden.case(name), layer_golden.manifest[name]
    x, w = Tensor(c["x"]), Tensor(c["w"])
    y = conv_nd(x, w, stride=m["stride"], padding=m["padding"], dilation=m["dilation"])
    o = max_pool(y, tuple(m["pool"]), m["pool_stride"]) if m["pool"] else y
    o.backward()
    dt = c["x"].dtype
    assert_close(o.numpy(), c["out"], dt, err_msg=name)
    assert_close(y.grad.get(), c["dconv"], dt, err_msg=name + " intermediate grad is kept")
    assert_close(x.grad.get(), c["dx"], dt, err_msg=name)
    assert_close(w.grad.get(), c["dw"], dt, err_msg=name)


# ---- seeded random cases against the oracle ---------------------------------------------

def _random_conv_case(rng, nd, dtype):
    W = tuple(int(rng.integers(1, 4)) for _ in range(nd))
    d = tuple(int(rng.integers(1, 3)) for _ in range(nd))
    s = tuple(int(rng.integers(1, 4)) for _ in range(nd))
    p = tuple(int(rng.integers(0, 3)) for _ in range(nd))
    G = tuple(int(rng.integers(1, 7 if nd < 3 else 4)) for _ in range(nd))
    X = tuple((g - 1) * si + (wi - 1) * di + 1 - 2 * pi for g, si, wi, di, pi in zip(G, s, W, d, p))
    if any(xi < 0 for xi in X):
        return None
    N, C, F = int(rng.integers(1, 4)), int(rng.integers(1, 6)), int(rng.integers(1, 6))
    x = rng.standard_normal((N, C, *X)).astype(dtype)
    w = rng.standard_normal((F, C, *W)).astype(dtype)
    return x, w, s, p, d


@pytest.mark.parametrize("nd", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_conv_random_vs_oracle(nd, dtype):
    # same space as tests/nnet/layers/test_conv.py:320-447 (1-3 spatial dims, random
    # stride / dilation / padding); analytic gradients from the oracle replace complex-step
    rng = np.random.default_rng(1000 + 10 * nd + (dtype == np.float32))
    done = 0
    while done < 25:
        case = _random_conv_case(rng, nd, dtype)
        if case is None:
            continue
        x, w, s, p, d = case
        done += 1
        xt, wt = Tensor(x), Tensor(w)
        y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
        want = oc.conv_fwd(x, w, s, p, d)
        msg = f"x{x.shape} w{w.shape} s{s} p{p} d{d}"
        assert y.shape == want.shape, msg
        assert_close(y.numpy(), want, dtype, err_msg=msg)
        g = rng.standard_normal(want.shape).astype(dtype)
        y.backward(g)
        assert_close(xt.grad.get(), oc.conv_dx(g, x.shape, w, s, p, d, dtype=dtype), dtype, err_msg=msg + " dx")
        assert_close(wt.grad.get(), oc.conv_dw(g, x, w.shape, s, p, d), dtype, err_msg=msg + " dw")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_conv_tile_boundaries_vs_oracle(dtype):
    """shapes chosen to straddle the 128/64/32-wide GEMM tiles and the K slabs"""
    rng = np.random.default_rng(5)
    for (N, C, F, X, W, s, p) in [(3, 7, 33, (11, 13), (3, 3), 1, 1), (2, 20, 65, (9, 9), (3, 3), 1, 1),
                                  (5, 9, 129, (7, 6), (2, 2), 1, 0), (2, 17, 130, (13, 13), (3, 3), 2, 1),
                                  (130, 3, 4, (5, 5), (3, 3), 1, 1)]:
        x = rng.standard_normal((N, C, *X)).astype(dtype)
        w = rng.standard_normal((F, C, *W)).astype(dtype)
        xt, wt = Tensor(x), Tensor(w)
        y = conv_nd(xt, wt, stride=s, padding=p)
        want = oc.conv_fwd(x, w, s, p, 1)
        assert_close(y.numpy(), want, dtype)
        g = rng.standard_normal(want.shape).astype(dtype)
        y.backward(g)
        assert_close(xt.grad.get(), oc.conv_dx(g, x.shape, w, s, p, 1, dtype=dtype), dtype)
        assert_close(wt.grad.get(), oc.conv_dw(g, x, w.shape, s, p, 1), dtype)


TC_CASES = [
    # N, C, F, X, W, stride, padding, dilation  (fp32, >= 16 channels in and out: tcgen05 path)
    (2, 32, 32, (8, 8), (1, 1), 1, 0, 1),
    (2, 32, 16, (9, 10), (3, 3), 1, 1, 1),
    (3, 16, 48, (12, 11), (3, 2), 1, (1, 0), 1),
    (2, 64, 128, (14, 14), (3, 3), 1, 1, 1),
    (2, 40, 72, (13, 13), (3, 3), 2, 1, 1),
    (2, 33, 130, (13, 13), (3, 3), 2, 0, 2),
    (1, 96, 200, (7, 9), (2, 3), (1, 2), (0, 1), (2, 1)),
    (3, 24, 20, (37,), (5,), 2, 2, 1),
    (2, 32, 64, (6, 7, 8), (3, 3, 3), 1, 1, 1),
    (2, 16, 16, (5, 9, 9), (2, 3, 3), (1, 2, 2), 0, 1),
    (5, 32, 32, (30, 30), (3, 3), 1, 1, 1),
    # 3-D with an unpadded outer axis: wgrad merges (batch, axis 0) into one TMA dimension
    (3, 32, 32, (7, 8, 8), (3, 3, 3), (2, 1, 1), (0, 1, 1), 1),
    (2, 32, 48, (10, 6, 6), (2, 3, 3), (2, 1, 1), (0, 1, 0), (3, 1, 1)),
    # stride-1 gathers with wide / dilated filters and ragged extents: the halo-reuse kernel's home turf
    (4, 32, 64, (20, 19), (5, 5), 1, 2, 1),
    (2, 64, 64, (17, 23), (3, 3), 1, (2, 1), (2, 3)),
    (3, 32, 32, (40,), (7,), 1, 3, 1),
    (17, 32, 160, (4, 5), (3, 3), 1, 1, 1),
    (2, 32, 32, (9, 20, 12), (3, 3, 3), 1, (1, 0, 1), 1),
    # an outer axis of extent 1 with three taps and padding 1 (found by scripts/fuzz_conv.py): X0 = G0 = 1 like a 2-D
    # layer, but the taps of a merged (batch, axis 0) TMA dimension would walk into the neighbouring sample
    (2, 24, 48, (1, 5, 9), (3, 2, 1), 1, (1, 2, 1), (1, 2, 1)),
    (3, 64, 128, (1, 12, 12), (3, 3, 3), 1, 1, 1),
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "x".join(map(str, c[:3])) + f"_{len(c[3])}d")
@pytest.mark.parametrize("path", ["halo", "halo2cta", "tma", "tma2cta", "cpasync", "mma"])
def test_conv_f32_paths_vs_oracle(case, path, monkeypatch):
    """every fp32 kernel family (tcgen05 fed by TMA with and without halo reuse, on one CTA and on CTA pairs
    (cta_group::2), tcgen05 fed by cp.async, mma.sync) against the oracle"""
    monkeypatch.setenv("MGB_CONV_F32_PATH", "tma" if path in ("tma2cta", "halo", "halo2cta") else path)
    monkeypatch.setenv("MGB_CONV_2CTA", "1" if path in ("tma2cta", "halo2cta") else "0")
    monkeypatch.setenv("MGB_CONV_HALO", "1" if path in ("halo", "halo2cta") else "0")   # "1": wherever the geometry allows
    N, C, F, X, W, s, p, d = case
    rng = np.random.default_rng(hash(case) % (2 ** 32))
    x = rng.standard_normal((N, C, *X)).astype(np.float32)
    w = rng.standard_normal((F, C, *W)).astype(np.float32)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    want = oc.conv_fwd(x.astype(np.float64), w.astype(np.float64), s, p, d)
    assert y.shape == want.shape
    assert_close(y.numpy(), want, np.float32, err_msg=f"{path} fwd {case}")
    g = rng.standard_normal(want.shape).astype(np.float32)
    y.backward(g)
    g64 = g.astype(np.float64)
    assert_close(xt.grad.get(), oc.conv_dx(g64, x.shape, w.astype(np.float64), s, p, d), np.float32,
                 err_msg=f"{path} dx {case}")
    assert_close(wt.grad.get(), oc.conv_dw(g64, x.astype(np.float64), w.shape, s, p, d), np.float32,
                 err_msg=f"{path} dw {case}")


F64_CASES = [
    # N, C, F, X, W, stride, padding, dilation  (float64; wide enough for the TMA-fed DMMA kernels)
    (2, 64, 128, (16, 16), (3, 3), 1, 1, 1),            # C2 in small: 128 x 128 tiles, 96-wide wgrad tiles
    (3, 32, 96, (14, 24), (3, 3), 1, 1, 1),             # ragged filter tile, ragged pixel boxes
    (2, 48, 64, (12, 16), (3, 3), 1, 1, 1),             # 64 filters: 256 x 64 tiles forward, wgrad stays on cp.async
    (2, 64, 128, (17, 16), (3, 3), (2, 1), (1, 1), 1),  # stride 2 on the middle axis (TMA traversal stride); odd rows
    (2, 64, 100, (12, 18), (2, 3), 1, (0, 2), (1, 2)),  # dilation 2 along the last axis: even shifts only
    (2, 32, 128, (20, 20), (5, 5), 1, 2, 1),            # 25 taps
    (2, 64, 128, (17, 16), (3, 2), 2, (1, 0), 1),       # stride 2 everywhere: forward falls back, dgrad classes use TMA
    (1, 32, 96, (5, 8, 12), (2, 3, 3), (1, 1, 1), (0, 1, 1), 1),   # 3-D
    (2, 64, 128, (15, 15), (3, 3), 1, 1, 1),            # odd row length: byte strides not multiples of 16, no TMA at all
    (2, 64, 128, (40,), (5,), 1, 2, 1),                 # 1-D
    (2, 64, 128, (17, 18), (3, 2), 2, (1, 0), 1),       # stride 2 on the innermost axis: contiguous rows read 2 apart
    (4, 32, 96, (37, 37), (5, 5), 2, 0, 2),             # config 3 in small: stride 2, dilation 2, odd rows of x and dy
]


@pytest.mark.parametrize("case", F64_CASES, ids=lambda c: "x".join(map(str, c[:3])) + f"_{len(c[3])}d")
@pytest.mark.parametrize("path", ["tma", "cpasync"])
def test_conv_f64_paths_vs_oracle(case, path, monkeypatch):
    """both float64 kernel families (DMMA fed by TMA with a producer warp, DMMA fed by cp.async) against the
    oracle at rtol 1e-10; shapes the TMA kernels decline run the cp.async kernels in both settings"""
    monkeypatch.setenv("MGB_F64_PATH", path)
    N, C, F, X, W, s, p, d = case
    rng = np.random.default_rng(hash(case) % (2 ** 32))
    x = rng.standard_normal((N, C, *X))
    w = rng.standard_normal((F, C, *W))
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    want = oc.conv_fwd(x, w, s, p, d)
    assert y.shape == want.shape
    assert_close(y.numpy(), want, np.float64, err_msg=f"{path} fwd {case}")
    g = rng.standard_normal(want.shape)
    y.backward(g)
    assert_close(xt.grad.get(), oc.conv_dx(g, x.shape, w, s, p, d), np.float64, err_msg=f"{path} dx {case}")
    assert_close(wt.grad.get(), oc.conv_dw(g, x, w.shape, s, p, d), np.float64, err_msg=f"{path} dw {case}")


WGRAD_ROWS_CASES = [
    # shapes for the row-halo wgrad kernel (conv_wgrad2.cuh): box shapes, c-block groups, dilation, strides, 3-D
    (4, 64, 128, (20, 20), (3, 3), 1, 1, 1),
    (3, 32, 64, (17, 19), (3, 3), 1, 1, 1),
    (2, 128, 256, (12, 12), (2, 2), 1, 0, 1),
    (2, 256, 96, (9, 9), (3, 3), (1, 2), 1, 1),
    (5, 16, 32, (30, 30), (5, 5), 1, 2, 2),
    (3, 64, 128, (9, 70), (3, 3), 1, 1, 1),
    (2, 32, 64, (6, 10, 10), (3, 3, 3), 1, 0, 1),
    (2, 64, 128, (6, 9, 12), (2, 3, 3), (2, 1, 1), (0, 1, 1), 1),
    (2, 64, 128, (40,), (5,), 1, 2, 1),
    # F <= 64: two column taps folded into the two halves of the 128 UMMA lanes (dy shifted by D2 columns)
    (3, 32, 64, (14, 15), (3, 2), 1, (1, 0), 1),        # two column taps: one CTA covers both
    (2, 32, 48, (12, 13), (3, 5), 1, (1, 2), 1),        # five column taps, 48 filters
    (2, 64, 64, (9, 21), (3, 3), 1, 1, (1, 3)),         # dilation 3 along the folded axis: shift by 3 columns
    (2, 32, 33, (5, 9, 11), (2, 3, 3), 1, (0, 1, 1), 1),  # 3-D, 33 filters
]


@pytest.mark.parametrize("case", WGRAD_ROWS_CASES, ids=lambda c: "x".join(map(str, c[:3])) + f"_{len(c[3])}d")
@pytest.mark.parametrize("rows", ["1", "0"])
def test_wgrad_row_halo_vs_oracle(case, rows, monkeypatch):
    """float32 dW from the row-halo kernel (forced wherever legal) and from the per-tap TMA kernel"""
    monkeypatch.setenv("MGB_WGRAD_ROWS", rows)
    N, C, F, X, W, s, p, d = case
    rng = np.random.default_rng(hash(case) % (2 ** 32))
    x = rng.standard_normal((N, C, *X)).astype(np.float32)
    w = rng.standard_normal((F, C, *W)).astype(np.float32)
    xt, wt = Tensor(x, constant=True), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    g = rng.standard_normal(y.shape).astype(np.float32)
    y.backward(g)
    want = oc.conv_dw(g.astype(np.float64), x.astype(np.float64), w.shape, s, p, d)
    assert_close(wt.grad.get(), want, np.float32, err_msg=f"rows={rows} dw {case}")


THIN_CASES = [
    # N, C, F, X, W, stride, padding, dilation: few input channels (C * taps <= 64) — conv_thin.cuh
    (4, 3, 32, (32, 32), (3, 3), 1, 1, 1),               # config 5, layer 1 in small
    (3, 3, 16, (20, 18), (3, 3), 1, 1, 1),               # config 1: 16 filters
    (2, 3, 70, (17, 23), (3, 3), (2, 1), (1, 0), 1),      # 70 filters: two filter groups; stride 2 on the rows
    (2, 1, 8, (30, 30), (5, 5), 1, 2, 2),                # one channel, dilation 2
    (2, 4, 24, (9, 40), (2, 7), (1, 3), (0, 3), 1),       # 56 taps x channels; stride 3 on the columns
    (3, 2, 40, (51,), (9,), 2, 4, 1),                    # 1-D
    (2, 3, 12, (6, 10, 9), (2, 3, 3), 1, (0, 1, 1), 1),   # 3-D: generic gather variants
    (2, 3, 32, (70, 70), (3, 3), 1, 1, 1),               # rows longer than one band pass; several bands per sample
]


@pytest.mark.parametrize("case", THIN_CASES, ids=lambda c: "x".join(map(str, c[:3])) + f"_{len(c[3])}d")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("path", ["mma", "fma", "off"])
def test_conv_thin_paths_vs_oracle(case, dtype, path, monkeypatch):
    """thin layers: band-staged mma.sync 3xTF32 kernels (float32), band-staged / gather FMA kernels (both dtypes),
    and the generic MMA tiles (MGB_CONV_THIN=0), forward + dX + dW against the oracle"""
    monkeypatch.setenv("MGB_CONV_THIN", "0" if path == "off" else "1")
    monkeypatch.setenv("MGB_THIN_MMA", "1" if path == "mma" else "0")
    N, C, F, X, W, s, p, d = case
    rng = np.random.default_rng(hash(case) % (2 ** 32))
    x = rng.standard_normal((N, C, *X)).astype(dtype)
    w = rng.standard_normal((F, C, *W)).astype(dtype)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=s, padding=p, dilation=d)
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    want = oc.conv_fwd(x64, w64, s, p, d)
    assert y.shape == want.shape
    assert_close(y.numpy(), want, dtype, err_msg=f"{path} fwd {case}")
    g = rng.standard_normal(want.shape).astype(dtype)
    y.backward(g)
    g64 = g.astype(np.float64)
    assert_close(xt.grad.get(), oc.conv_dx(g64, x.shape, w64, s, p, d), dtype, err_msg=f"{path} dx {case}")
    assert_close(wt.grad.get(), oc.conv_dw(g64, x64, w.shape, s, p, d), dtype, err_msg=f"{path} dw {case}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_pool_random_vs_oracle_bit_exact(dtype):
    rng = np.random.default_rng(77)
    for _ in range(40):
        nd = int(rng.integers(1, 4))
        pool = tuple(int(rng.integers(1, 4)) for _ in range(nd))
        stride = tuple(int(rng.integers(1, 4)) for _ in range(nd))
        G = tuple(int(rng.integers(1, 6)) for _ in range(nd))
        X = tuple((g - 1) * s + p for g, s, p in zip(G, stride, pool))
        lead = tuple(int(rng.integers(1, 4)) for _ in range(int(rng.integers(0, 3))))
        # small alphabet half of the time: ties exercise the first-max rule
        if rng.integers(0, 2):
            x = rng.integers(-2, 3, size=(*lead, *X)).astype(dtype)
        else:
            x = rng.standard_normal((*lead, *X)).astype(dtype)
        xt = Tensor(x)
        y = max_pool(xt, pool, stride)
        np.testing.assert_array_equal(y.numpy(), oc.max_pool_fwd(x, pool, stride))
        g = rng.standard_normal(y.shape).astype(dtype)
        y.backward(g)
        np.testing.assert_array_equal(xt.grad.get(), oc.max_pool_bwd(g, x, pool, stride).astype(dtype))


# ---- BASELINE.json full sizes: properties + spot checks --------------------------------

def _full_config(name):
    if name == "C1":
        return dict(x=(32, 3, 32, 32), w=(16, 3, 3, 3), s=1, p=1, d=1, pool=(2, 2), dt=np.float64)
    if name == "C2":
        return dict(x=(256, 64, 56, 56), w=(128, 64, 3, 3), s=1, p=1, d=1, pool=(2, 2), dt=np.float32)
    if name == "C2f64":
        return dict(x=(64, 64, 56, 56), w=(128, 64, 3, 3), s=1, p=1, d=1, pool=(2, 2), dt=np.float64)
    if name == "C3":   # as written it is invalid; 65x65 is the nearest valid shape
        return dict(x=(16, 128, 65, 65), w=(256, 128, 5, 5), s=2, p=0, d=2, pool=None, dt=np.float64)
    if name == "C4":   # the per-GPU shard at 8 GPUs: 4 samples
        return dict(x=(4, 32, 32, 64, 64), w=(64, 32, 3, 3, 3), s=1, p=0, d=1, pool=(2, 2, 2), dt=np.float32)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["C1", "C2", "C2f64", "C3", "C4"])
def test_full_size_spot_checks_and_properties(name):
    cfg = _full_config(name)
    dt = cfg["dt"]
    rng = np.random.default_rng(0)
    x = rng.standard_normal(cfg["x"]).astype(dt)
    w = np.random.default_rng(1).standard_normal(cfg["w"]).astype(dt)
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=cfg["s"], padding=cfg["p"], dilation=cfg["d"])
    out = max_pool(y, cfg["pool"], 2) if cfg["pool"] else y
    out.backward()
    yh, dxh, dwh = y.numpy(), xt.grad.get(), wt.grad.get()

    # (1) spot check: two whole samples against the oracle (forward, pooling, dX)
    for n in (0, cfg["x"][0] - 1):
        xs = x[n:n + 1]
        y_ref = oc.conv_fwd(xs, w, cfg["s"], cfg["p"], cfg["d"])
        assert_close(yh[n:n + 1], y_ref, dt, err_msg=f"{name} fwd sample {n}")
        if cfg["pool"]:
            # pooling is exact on whatever y the GPU produced
            np.testing.assert_array_equal(out.numpy()[n:n + 1], oc.max_pool_fwd(yh[n:n + 1], cfg["pool"], 2))
            dy = oc.max_pool_bwd(np.ones(out.shape[1:], dt)[None], yh[n:n + 1], cfg["pool"], 2).astype(dt)
            np.testing.assert_array_equal(y.grad.get()[n:n + 1], dy)
        else:
            dy = np.ones_like(y_ref)
        dx_ref = oc.conv_dx(dy, xs.shape, w, cfg["s"], cfg["p"], cfg["d"], dtype=dt)
        assert_close(dxh[n:n + 1], dx_ref, dt, err_msg=f"{name} dx sample {n}")

    # (2) dW checksum: sum_n of per-sample oracle dW over a few samples, compared through
    #     linearity: dW(full) - dW(without those samples) == dW(those samples)
    keep = cfg["x"][0] // 2
    x2, w2 = Tensor(x[:keep]), Tensor(w)
    y2 = conv_nd(x2, w2, stride=cfg["s"], padding=cfg["p"], dilation=cfg["d"])
    (max_pool(y2, cfg["pool"], 2) if cfg["pool"] else y2).backward()
    x3, w3 = Tensor(x[keep:]), Tensor(w)
    y3 = conv_nd(x3, w3, stride=cfg["s"], padding=cfg["p"], dilation=cfg["d"])
    (max_pool(y3, cfg["pool"], 2) if cfg["pool"] else y3).backward()
    assert_close(w2.grad.get() + w3.grad.get(), dwh, dt, err_msg=f"{name}: dW is additive over the batch")
    np.testing.assert_array_equal(x2.grad.get(), dxh[:keep])   # dX is per-sample: bit-identical shards

    # (3) adjoint identity <conv(x), g> == <x, dX(g)> == <w, dW(g)> for the all-ones-through-pool g
    gy = y.grad.get().astype(np.float64)
    lhs = float((yh.astype(np.float64) * gy).sum())
    rtol = 1e-9 if dt == np.float64 else 2e-4
    assert abs(lhs - float((x.astype(np.float64) * dxh).sum())) <= rtol * max(1.0, abs(lhs)), name
    assert abs(lhs - float((w.astype(np.float64) * dwh).sum())) <= rtol * max(1.0, abs(lhs)), name

    # (5) dW / dX of the FULL-SIZE layer against the oracle on a 2-sample slice (full spatial extent and channel
    #     counts: the same kernels, tiles and split-K plan shape as the whole batch; (2) extends it by additivity).
    #     The oracle is fed the dy the GPU routed (bit-exact on the GPU's own y, checked in (1)): in float32 a
    #     near-tie inside a pooling window may legitimately resolve differently on y values that differ by 1e-6
    x5, w5 = Tensor(x[:2]), Tensor(w)
    y5 = conv_nd(x5, w5, stride=cfg["s"], padding=cfg["p"], dilation=cfg["d"])
    (max_pool(y5, cfg["pool"], 2) if cfg["pool"] else y5).backward()
    dy5 = y5.grad.get()
    if cfg["pool"]:
        np.testing.assert_array_equal(
            dy5, oc.max_pool_bwd(np.ones((2, *out.shape[1:]), dt), y5.numpy(), cfg["pool"], 2).astype(dt))
    dw_ref5 = oc.conv_dw(dy5, x[:2], w.shape, cfg["s"], cfg["p"], cfg["d"]).astype(dt)
    dx_ref5 = oc.conv_dx(dy5, x[:2].shape, w, cfg["s"], cfg["p"], cfg["d"], dtype=dt)
    assert_close(w5.grad.get(), dw_ref5, dt, err_msg=f"{name}: dW of a 2-sample slice vs the oracle")
    assert_close(x5.grad.get(), dx_ref5, dt, err_msg=f"{name}: dX of a 2-sample slice vs the oracle")

    # (4) strided+dilated dX has exact zeros where no window lands (config 3)
    if name == "C3":
        assert np.all(dxh[..., 1::2, :] == 0) and np.all(dxh[..., :, 1::2] == 0)
