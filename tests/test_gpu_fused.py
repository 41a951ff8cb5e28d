"""The opt-in fused Operation ``conv_pool`` against the unfused drop-in pair
``max_pool(conv_nd(...))``: pooled values, first-max routing and both gradients must agree BIT FOR BIT
(same accumulator values, same scan order), for the geometries the fused kernels cover and — through the
op's internal fallback — for the ones they do not.  The unfused pair itself is pinned to the reference
in test_gpu_parity.py."""
import numpy as np
import pytest

import mygrad_b200 as mg
from mygrad_b200 import Tensor
from mygrad_b200.nnet.layers import conv_nd, conv_pool, max_pool
from oracle import conv_pool_oracle as oc

from conftest import GoldenFile, assert_close

pytestmark = pytest.mark.gpu


def _pair(x, w, kw, pool, ps, g):
    xt, wt = Tensor(x), Tensor(w)
    out = max_pool(conv_nd(xt, wt, **kw), pool, ps)
    out.backward(g)
    return out.numpy(), xt.grad.get(), wt.grad.get()


def _fused(x, w, kw, pool, ps, g):
    xt, wt = Tensor(x), Tensor(w)
    out = conv_pool(xt, wt, pool=pool, pool_stride=ps, **kw)
    op = out.creator
    out.backward(g)
    return out.numpy(), xt.grad.get(), wt.grad.get(), op.fused


FUSED_CASES = [
    # name, x shape, w shape, padding, dilation
    ("small", (2, 16, 8, 8), (16, 16, 3, 3), 1, 1),
    ("odd_batch_ragged", (3, 32, 12, 20), (48, 32, 3, 3), 1, 1),
    ("one_sample", (1, 16, 16, 16), (32, 16, 3, 3), 1, 1),
    ("wide_n_tile", (2, 32, 10, 10), (160, 32, 3, 3), 1, 1),          # two channel tiles
    ("filter5_pad2", (2, 16, 12, 12), (16, 16, 5, 5), 2, 1),
    ("dilated", (2, 16, 14, 14), (24, 16, 3, 3), 2, 2),
    ("filter1x1", (4, 64, 8, 8), (32, 64, 1, 1), 0, 1),
    ("no_pad", (2, 16, 10, 18), (16, 16, 3, 3), 0, 1),
    ("c2_geometry", (5, 64, 56, 56), (128, 64, 3, 3), 1, 1),
    ("channel_padding", (2, 20, 8, 8), (40, 20, 3, 3), 1, 1),          # 20 -> 32, 40 -> 48 padded channels
]


@pytest.mark.parametrize("name,xs,ws,pad,dil", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_equals_unfused_bitwise(name, xs, ws, pad, dil, monkeypatch):
    # the fused forward always runs the halo kernel; the unfused conv_nd picks halo or the per-tap TMA kernel
    # by a cost model (they group the 3xTF32 partial sums differently).  Pin the unfused side to the halo
    # kernel — its default at layer sizes like C2 — so that "same kernel, same bits" is what is compared.
    monkeypatch.setenv("MGB_CONV_HALO", "1")
    rng = np.random.default_rng(sum(map(ord, name)))
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    kw = dict(stride=1, padding=pad, dilation=dil)
    y_sp = oc.conv_out_shape(xs[2:], ws[2:], (1, 1), (pad, pad), (dil, dil))
    g = rng.standard_normal((xs[0], ws[0], y_sp[0] // 2, y_sp[1] // 2)).astype(np.float32)
    want = _pair(x, w, kw, (2, 2), 2, g)
    got = _fused(x, w, kw, (2, 2), 2, g)
    assert got[3], f"{name}: expected the fused kernels to apply"
    for nm, a, b in zip(("out", "dx", "dw"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=f"{name} {nm}")
    # and against the oracle within the north-star tolerance (fed the GPU-routed dy, see test_gpu_parity (5))
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, **kw)
    o = max_pool(y, (2, 2), 2)
    o.backward(g)
    dy = y.grad.get()
    assert_close(got[1], oc.conv_dx(dy, xs, w, 1, pad, dil, dtype=np.float32), np.float32, err_msg=name + " dx")
    assert_close(got[2], oc.conv_dw(dy, x, ws, 1, pad, dil).astype(np.float32), np.float32, err_msg=name + " dw")
    assert_close(got[0], oc.max_pool_fwd(oc.conv_fwd(x, w, 1, pad, dil), (2, 2), 2), np.float32, err_msg=name)


def test_fused_ties_and_nan_route_like_the_reference(monkeypatch):
    monkeypatch.setenv("MGB_CONV_HALO", "1")
    _ties_and_nan()


def test_fused_within_tolerance_of_unfused_whatever_kernel_it_picks():
    rng = np.random.default_rng(21)
    for xs, ws in (((2, 32, 10, 10), (160, 32, 3, 3)), ((3, 16, 6, 6), (16, 16, 3, 3))):
        x, w = rng.standard_normal(xs).astype(np.float32), rng.standard_normal(ws).astype(np.float32)
        g = rng.standard_normal((xs[0], ws[0], xs[2] // 2, xs[3] // 2)).astype(np.float32)
        kw = dict(stride=1, padding=1, dilation=1)
        want, got = _pair(x, w, kw, (2, 2), 2, g), _fused(x, w, kw, (2, 2), 2, g)
        for nm, a, b in zip(("out", "dx", "dw"), got, want):
            assert_close(a, b, np.float32, err_msg=nm)


def _ties_and_nan():
    """integer data makes exact ties inside windows (first maximum in row-major order wins); a NaN in x makes
    whole neighbourhoods of y NaN (NaN counts as maximal, the first NaN wins)"""
    rng = np.random.default_rng(5)
    x = rng.integers(-1, 2, size=(2, 16, 12, 12)).astype(np.float32)
    w = rng.integers(-1, 2, size=(16, 16, 3, 3)).astype(np.float32)
    g = rng.standard_normal((2, 16, 6, 6)).astype(np.float32)
    kw = dict(stride=1, padding=1, dilation=1)
    want, got = _pair(x, w, kw, (2, 2), 2, g), _fused(x, w, kw, (2, 2), 2, g)
    assert got[3]
    y = oc.conv_fwd(x, w, 1, 1, 1)
    ties = (np.sort(y.reshape(2, 16, 6, 2, 6, 2).transpose(0, 1, 2, 4, 3, 5).reshape(2, 16, 6, 6, 4), -1)[..., -1:]
            == y.reshape(2, 16, 6, 2, 6, 2).transpose(0, 1, 2, 4, 3, 5).reshape(2, 16, 6, 6, 4)).sum(-1)
    assert (ties > 1).any(), "the case is meant to contain tied windows"
    for nm, a, b in zip(("out", "dx", "dw"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=nm)
    # integer arithmetic is exact in float32 here: the pooled values and routing equal the reference algorithm's
    np.testing.assert_array_equal(got[0], oc.max_pool_fwd(y, (2, 2), 2))
    dy = oc.max_pool_bwd(g, y, (2, 2), 2).astype(np.float32)
    assert_close(got[1], oc.conv_dx(dy, x.shape, w, 1, 1, 1, dtype=np.float32), np.float32, err_msg="tie routing")
    assert_close(got[2], oc.conv_dw(dy, x, w.shape, 1, 1, 1).astype(np.float32), np.float32, err_msg="tie routing dw")

    x[0, 3, 5, 5] = np.nan
    x[1, 0, 0, 0] = np.nan
    want, got = _pair(x, w, kw, (2, 2), 2, g), _fused(x, w, kw, (2, 2), 2, g)
    assert np.isnan(got[0]).any()
    for nm, a, b in zip(("out", "dx", "dw"), got, want):
        np.testing.assert_array_equal(a, b, err_msg="nan " + nm)


UNFUSED_CASES = [
    # geometries outside the fused kernels: the op must still be correct (internal fallback)
    ("f64", (2, 16, 8, 8), (16, 16, 3, 3), dict(stride=1, padding=1), (2, 2), 2, np.float64),
    ("conv_stride2", (2, 16, 9, 9), (16, 16, 3, 3), dict(stride=2, padding=1), (1, 1), 1, np.float32),
    ("pool3", (2, 16, 9, 9), (16, 16, 3, 3), dict(stride=1, padding=1), (3, 3), 3, np.float32),
    ("pool_overlap", (2, 16, 8, 8), (16, 16, 3, 3), dict(stride=1, padding=1), (2, 2), 1, np.float32),
    ("few_channels", (2, 3, 8, 8), (4, 3, 3, 3), dict(stride=1, padding=1), (2, 2), 2, np.float32),
    ("conv1d", (2, 16, 16), (16, 16, 3), dict(stride=1, padding=1), (2,), 2, np.float32),
    ("conv3d", (1, 16, 4, 6, 6), (16, 16, 3, 3, 3), dict(stride=1, padding=1), (2, 2, 2), 2, np.float32),
    ("pool_trailing_axis_only", (2, 16, 8, 8), (16, 16, 3, 3), dict(stride=1, padding=1), (2,), 2, np.float32),
    ("odd_output", (2, 16, 9, 9), (16, 16, 3, 3), dict(stride=1, padding=0), (7, 7), 7, np.float32),
]


@pytest.mark.parametrize("name,xs,ws,kw,pool,ps,dt", UNFUSED_CASES, ids=[c[0] for c in UNFUSED_CASES])
def test_unsupported_geometries_fall_back_inside_the_op(name, xs, ws, kw, pool, ps, dt):
    rng = np.random.default_rng(11)
    x, w = rng.standard_normal(xs).astype(dt), rng.standard_normal(ws).astype(dt)
    xt, wt = Tensor(x), Tensor(w)
    out = max_pool(conv_nd(xt, wt, **kw), pool, ps)
    g = rng.standard_normal(out.shape).astype(dt)
    want = _pair(x, w, kw, pool, ps, g)
    got = _fused(x, w, kw, pool, ps, g)
    assert not got[3], f"{name}: not a geometry of the fused kernels"
    for nm, a, b in zip(("out", "dx", "dw"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=f"{name} {nm}")


@pytest.mark.parametrize("name", GoldenFile("layer_cases.npz").names())
def test_conv_pool_vs_reference_layer_goldens(name, layer_golden):
    """whole layers run by the LIVE reference (tests/golden/make_golden.py): conv_pool must give the same
    pooled output, x.grad and w.grad as the reference's max_pool(conv_nd(...)).backward()"""
    c, m = layer_golden.case(name), layer_golden.manifest[name]
    if m.get("pool") is None:
        pytest.skip("layer without pooling")
    x, w = Tensor(c["x"]), Tensor(c["w"])
    out = conv_pool(x, w, stride=m["stride"], padding=m["padding"], dilation=m["dilation"],
                    pool=tuple(m["pool"]), pool_stride=m["pool_stride"])
    out.backward()
    assert_close(out.numpy(), c["out"], c["out"].dtype, err_msg=name)
    assert_close(x.grad.get(), c["dx"], c["dx"].dtype, err_msg=name + " dx")
    assert_close(w.grad.get(), c["dw"], c["dw"].dtype, err_msg=name + " dw")


def test_conv_pool_graph_semantics_and_errors():
    rng = np.random.default_rng(3)
    x = Tensor(rng.standard_normal((2, 16, 8, 8)).astype(np.float32))
    w = Tensor(rng.standard_normal((16, 16, 3, 3)).astype(np.float32))
    out = conv_pool(x, w, stride=1, padding=1, pool=(2, 2), pool_stride=2)
    assert out.shape == (2, 16, 4, 4) and out.dtype == np.float32 and not out.constant
    assert out.creator.fused and tuple(out.creator.pool) == (2, 2) and tuple(out.creator.stride) == (1, 1)
    out.backward()
    assert x.grad.shape == x.shape and w.grad.shape == w.shape and out.creator is None
    # constants propagate as for any Operation
    xc = Tensor(x.data, constant=True)
    o2 = conv_pool(xc, w, stride=1, padding=1, pool=(2, 2), pool_stride=2)
    o2.backward()
    assert xc.grad is None and w.grad is not None
    assert conv_pool(xc, Tensor(w.data, constant=True), stride=1, padding=1, pool=(2, 2), pool_stride=2).constant
    # the reference's exception types (conv.py:41-74, 373-387; pooling.py:52-86)
    with pytest.raises(ValueError):
        conv_pool(np.zeros((2, 16, 8), np.float32), w, stride=1, pool=(2,), pool_stride=2)
    with pytest.raises(ValueError):
        conv_pool(x, w, stride=3, padding=0, pool=(2, 2), pool_stride=2)     # non-integer conv output
    with pytest.raises(ValueError):
        conv_pool(x, w, stride=1, padding=1, pool=(3, 3), pool_stride=2)     # non-integer pooled output
    with pytest.raises(AssertionError):
        conv_pool(x, w, stride=0, padding=1, pool=(2, 2), pool_stride=2)
    with pytest.raises(AssertionError):
        conv_pool(x, w, stride=1, padding=1, pool=(2, 0), pool_stride=2)


def test_c_abi_fused_entry_points_reject_unsupported_geometry():
    import ctypes as C

    from mygrad_b200._lib import ConvDesc, PoolDesc, lib

    cd, pd = ConvDesc(), PoolDesc()
    cd.dtype, cd.nd, cd.N, cd.C, cd.F = 1, 2, 2, 16, 16          # float64: not covered
    for i in range(2):
        cd.X[i], cd.W[i], cd.stride[i], cd.pad[i], cd.dil[i] = 8, 3, 1, 1, 1
        pd.X[i], pd.P[i], pd.stride[i] = 8, 2, 2
    pd.dtype, pd.nd, pd.mode, pd.lead = 1, 2, 0, 32
    yes = C.c_int(7)
    assert lib.mgb_conv_pool_supported(C.byref(cd), C.byref(pd), C.byref(yes)) == 0 and yes.value == 0
    assert lib.mgb_conv_pool_fwd(C.byref(cd), C.byref(pd), None, None, None, None, None, None, 0, None) != 0
    assert b"not supported" in lib.mgb_last_error()


def test_unfused_pair_pool_backward_feeds_the_conv_planes(monkeypatch):
    """max_pool(conv_nd(...)).backward(): when the pooled tensor comes from a tensor-core convolution and has no
    other consumer, MaxPoolND.backward_var writes dy AND the convolution's dy planes in one kernel.  y.grad, x.grad
    and w.grad must be bit-identical to the two-kernel path (mgb_pool_bwd + mgb_conv_make_planes)."""
    rng = np.random.default_rng(17)
    x = rng.standard_normal((3, 32, 12, 20)).astype(np.float32)
    w = rng.standard_normal((48, 32, 3, 3)).astype(np.float32)
    g = rng.standard_normal((3, 48, 6, 10)).astype(np.float32)

    def run():
        xt, wt = Tensor(x), Tensor(w)
        y = conv_nd(xt, wt, stride=1, padding=1)
        o = max_pool(y, (2, 2), 2)
        conv_op = y.creator
        o.backward(g)
        return y.grad.get(), xt.grad.get(), wt.grad.get(), conv_op

    monkeypatch.setenv("MGB_POOL_FEEDS_CONV", "0")
    want = run()
    monkeypatch.setenv("MGB_POOL_FEEDS_CONV", "1")
    n0 = C_launches()
    got = run()
    one_pass = C_launches() - n0
    monkeypatch.setenv("MGB_POOL_FEEDS_CONV", "0")
    n0 = C_launches()
    run()
    two_pass = C_launches() - n0
    assert one_pass == two_pass - 1, (one_pass, two_pass)        # plane prep of dy folded into the pool backward
    for nm, a, b in zip(("dy", "dx", "dw"), got, want):
        np.testing.assert_array_equal(a, b, err_msg=nm)
    yh = oc.conv_fwd(x, w, 1, 1, 1)
    # x constant: only dW runs and must still find the planes
    monkeypatch.setenv("MGB_POOL_FEEDS_CONV", "1")
    xt, wt = Tensor(x, constant=True), Tensor(w)
    max_pool(conv_nd(xt, wt, stride=1, padding=1), (2, 2), 2).backward(g)
    np.testing.assert_array_equal(wt.grad.get(), want[2])
    # a second consumer of y: grads accumulate into y.grad, so the planes of the pool's contribution alone
    # would be stale — the one-pass path must stay out of the way
    xt, wt = Tensor(x), Tensor(w)
    y = conv_nd(xt, wt, stride=1, padding=1)
    o1, o2 = max_pool(y, (2, 2), 2), max_pool(y, (2, 2), 2)
    (o1 + o2).backward(g)
    np.testing.assert_array_equal(y.grad.get(), 2 * want[0])
    assert_close(xt.grad.get(), 2 * want[1], np.float32)
    assert_close(wt.grad.get(), 2 * want[2], np.float32)


def C_launches() -> int:
    import ctypes as C

    from mygrad_b200._lib import lib

    n = C.c_uint64(0)
    lib.mgb_launch_count(C.byref(n))
    return int(n.value)
