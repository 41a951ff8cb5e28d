"""Tensor.backward / Operation.backward semantics on device buffers, and the raw C ABI
driven with foreign (torch) buffers through __cuda_array_interface__."""
import ctypes as C

import numpy as np
import pytest

import mygrad_b200 as mg
from conftest import assert_close
from mygrad_b200 import DeviceArray, InvalidBackprop, InvalidGradient, Operation, Tensor
from mygrad_b200._lib import ConvDesc, PoolDesc, check, lib
from mygrad_b200.nnet.layers import conv_nd, max_pool
from oracle import conv_pool_oracle as oc

pytestmark = pytest.mark.gpu


class Add(Operation):
    """a + b with NumPy broadcasting; backward_var hands `grad` itself back, exactly like
    mygrad's Add (math/arithmetic/ops.py:28-32) — Operation.backward must copy and un-broadcast."""

    def __call__(self, a, b):
        self.variables = (a, b)
        out = (a.data.get().astype(np.result_type(a.dtype, b.dtype)) + b.data.get())
        return mg.asdevice(out)

    def backward_var(self, grad, index, **kwargs):
        return grad


def add(a, b):
    return Tensor._op(Add, a, b)


def test_multi_consumer_accumulation_vs_golden(graph_golden):
    c = graph_golden.case("acc")
    x, w1, w2 = Tensor(c["x"]), Tensor(c["w1"]), Tensor(c["w2"])
    total = add(conv_nd(x, w1, stride=1, padding=1), conv_nd(x, w2, stride=1, padding=1))
    o = max_pool(total, (2, 2), 2)
    o.backward()
    assert_close(o.numpy(), c["out"])
    assert_close(total.grad.get(), c["dtotal"])
    assert_close(x.grad.get(), c["dx"])          # two arrivals: first copy, then `+=`
    assert_close(w1.grad.get(), c["dw1"])
    assert_close(w2.grad.get(), c["dw2"])


def test_bias_unbroadcast_vs_golden(graph_golden):
    c, a = graph_golden.case("bias"), graph_golden.case("acc")
    x, w1 = Tensor(a["x"]), Tensor(a["w1"])
    b4, b3 = Tensor(c["b4"]), Tensor(c["b3"])
    y = add(add(conv_nd(x, w1, stride=1, padding=1), b4), b3)
    y.backward(c["g"])
    assert_close(y.numpy(), c["y"])
    assert b4.grad.shape == (1, 3, 1, 1) and b3.grad.shape == (3, 1, 1)
    assert_close(b4.grad.get(), c["db4"])
    assert_close(b3.grad.get(), c["db3"])
    assert_close(x.grad.get(), c["dx"])


def test_reduce_broadcast_shapes_vs_golden(graph_golden):
    n = int(graph_golden.npz["rb_count"])
    for i in range(n):
        c = graph_golden.case(f"rb{i}")
        vshape = tuple(int(v) for v in c["vshape"])
        for dt in (np.float64, np.float32):
            out = Operation.grad_post_process_fn(mg.asdevice(c["g"].astype(dt)), vshape)
            assert out.shape == vshape
            assert_close(out.get(), c["out"].astype(dt), dt, err_msg=f"{c['g'].shape}->{vshape}")
    with pytest.raises(ValueError):
        Operation.grad_post_process_fn(mg.asdevice(np.ones((3,))), (2, 3))


def test_reduce_broadcast_large_bias_shape():
    # config-5-like bias grad (N,F,H,W) -> (1,F,1,1): many rows per output, split across CTAs
    rng = np.random.default_rng(3)
    g = rng.standard_normal((64, 32, 32, 32)).astype(np.float32)
    out = mg.asdevice(g).sum_to_shape((1, 32, 1, 1)).get()
    want = g.astype(np.float64).sum(axis=(0, 2, 3), keepdims=True)
    np.testing.assert_allclose(out, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    out2 = mg.asdevice(g.reshape(64 * 32, 1024)).sum_to_shape((1024,)).get()
    np.testing.assert_allclose(out2, g.reshape(-1, 1024).astype(np.float64).sum(0), rtol=1e-5, atol=1e-3)


def test_pool_grad_dtype_follows_x(graph_golden):
    c = graph_golden.case("pool32")
    x = Tensor(c["x"])
    o = max_pool(x, (3, 3), 2)
    o.backward(c["g"])
    assert x.grad.dtype == np.float32            # float64 scatter buffer, cast back (operation_base.py:223-224)
    np.testing.assert_array_equal(x.grad.get(), c["dx"])


def test_backward_seed_and_graph_clearing():
    # tests/tensor_base/test_backward.py:21-96, test_no_null_grad_semantics.py:19-36,
    # test_backprop_through_non_scalar.py:11-18
    rng = np.random.default_rng(0)
    xh, wh = rng.standard_normal((2, 2, 6, 6)), rng.standard_normal((3, 2, 3, 3))
    x, w = Tensor(xh), Tensor(wh)
    y = conv_nd(x, w, stride=1)
    y.backward()                                  # non-scalar terminal == sum().backward()
    assert y.grad.dtype == y.dtype
    np.testing.assert_array_equal(y.grad.get(), np.ones(y.shape))
    assert_close(w.grad.get(), oc.conv_dw(np.ones(y.shape), xh, wh.shape, 1))
    assert y.creator is None and not x._ops and not w._ops
    first = x.grad
    y2 = conv_nd(x, w, stride=1)                  # re-use nulls the grads of the inputs
    assert x.grad is None and w.grad is None and first is not None
    y2.backward(2.0)                              # scalar grad broadcasts
    np.testing.assert_array_equal(y2.grad.get(), np.full(y2.shape, 2.0))
    assert_close(x.grad.get(), 2 * oc.conv_dx(np.ones(y.shape), xh.shape, wh, 1))
    y3 = conv_nd(x, w, stride=1)
    with pytest.raises(ValueError):
        y3.backward(np.ones((5,)))                # not broadcast-compatible
    y3.clear_graph()
    xf = Tensor(xh.astype(np.float32))
    y4 = conv_nd(xf, w, stride=1)                 # f32 x, f64 w -> f64 y, x.grad f32, w.grad f64
    y4.backward(np.ones(y4.shape, dtype=np.float32))
    assert (y4.dtype, xf.grad.dtype, w.grad.dtype) == (np.float64, np.float32, np.float64)


def test_constants_and_no_autodiff():
    rng = np.random.default_rng(1)
    xh, wh = rng.standard_normal((1, 2, 5, 5)), rng.standard_normal((2, 2, 2, 2))
    x, w = Tensor(xh, constant=True), Tensor(wh)
    y = conv_nd(x, w, stride=1)
    assert y.constant is False
    y.backward()
    assert x.grad is None and w.grad is not None  # constants are skipped (operation_base.py:179)
    yc = conv_nd(xh, wh, stride=1)                # bare arrays become constant tensors
    assert yc.constant is True
    yc.backward()
    assert yc.grad is None
    with mg.no_autodiff:
        yn = conv_nd(Tensor(xh), w, stride=1)
        assert yn.creator is None
        yn.backward()
        assert yn.grad is None
    assert_close(yn.numpy(), oc.conv_fwd(xh, wh, 1))


def test_invalid_backprop_and_invalid_gradient():
    class Bad(Operation):
        def __call__(self, a):
            self.variables = (a,)
            return a.data.copy()

        def backward_var(self, grad, index, **kwargs):
            return "not a gradient"

    x = Tensor(np.ones((2, 2)))
    with pytest.raises(InvalidGradient):
        Tensor._op(Bad, x).backward()             # tests/test_operation_base.py:15-57

    x = Tensor(np.ones((1, 1, 4, 4)))
    y = conv_nd(x, np.ones((1, 1, 2, 2)), stride=1)
    z = max_pool(y, (3, 3), 1)
    y.clear_graph()                               # z's creator still points at the cleared y
    with pytest.raises(InvalidBackprop):
        z.backward()


# ---- raw C ABI with foreign buffers ---------------------------------------------------------

def test_c_abi_with_torch_buffers():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    xh = rng.standard_normal((3, 4, 9, 9)).astype(np.float32)
    wh = rng.standard_normal((5, 4, 3, 3)).astype(np.float32)
    xt, wt = torch.from_numpy(xh).to(dev), torch.from_numpy(wh).to(dev)
    yt = torch.empty((3, 5, 9, 9), device=dev)
    d = ConvDesc()
    d.dtype, d.nd, d.N, d.C, d.F = 0, 2, 3, 4, 5
    for i in range(2):
        d.X[i], d.W[i], d.stride[i], d.pad[i], d.dil[i] = 9, 3, 1, 1, 1
    n = C.c_size_t(0)
    check(lib.mgb_conv_workspace_bytes(C.byref(d), 0, C.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    check(lib.mgb_conv_fwd(C.byref(d), C.c_void_p(xt.data_ptr()), C.c_void_p(wt.data_ptr()),
                           C.c_void_p(yt.data_ptr()), C.c_void_p(ws.data_ptr()), n, None))
    check(lib.mgb_device_synchronize())
    assert_close(yt.cpu().numpy(), oc.conv_fwd(xh, wh, 1, 1, 1), np.float32)
    # too-small workspace is an error status, not a crash
    assert lib.mgb_conv_fwd(C.byref(d), C.c_void_p(xt.data_ptr()), C.c_void_p(wt.data_ptr()),
                            C.c_void_p(yt.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(8), None) != 0
    assert b"workspace" in lib.mgb_last_error()
    # zero-copy import of a torch tensor through __cuda_array_interface__, and export back
    t = Tensor(xt, copy=False)
    assert t.data.ptr == xt.data_ptr()
    out = max_pool(t, (3, 3), 3)
    back = torch.as_tensor(out.data, device=dev)
    np.testing.assert_array_equal(back.cpu().numpy(), oc.max_pool_fwd(xh, (3, 3), 3))
    # torch's own conv as an independent third opinion (not part of the reference)
    ref = torch.nn.functional.conv2d(xt.double(), wt.double(), padding=1).float().cpu().numpy()
    assert_close(yt.cpu().numpy(), ref, np.float32)


def test_pool_c_abi_rejects_bad_descriptor():
    d = PoolDesc()
    d.dtype, d.nd, d.mode, d.lead = 0, 2, 0, 1
    d.X[0], d.X[1], d.P[0], d.P[1], d.stride[0], d.stride[1] = 5, 5, 2, 2, 2, 2
    buf = mg.empty((25,), np.float32)
    assert lib.mgb_pool_fwd(C.byref(d), C.c_void_p(buf.ptr), C.c_void_p(buf.ptr), None) != 0
    assert b"incompatible" in lib.mgb_last_error()


def test_launch_counter_counts_kernels():
    before, after = C.c_uint64(0), C.c_uint64(0)
    lib.mgb_launch_count(C.byref(before))
    mg.full((1024,), 1.0, np.float32)
    lib.mgb_launch_count(C.byref(after))
    assert after.value == before.value + 1


def test_host_pipeline_matches_single_pass():
    """chunked H2D/compute/D2H pipeline == one pass over the whole batch (dW is additive)"""
    rng = np.random.default_rng(9)
    for dt, chunks in ((np.float32, 3), (np.float64, 4), (np.float32, 16)):
        xh = rng.standard_normal((10, 16, 12, 12)).astype(dt)
        wh = rng.standard_normal((24, 16, 3, 3)).astype(dt)
        px, pw = mg.pinned_empty(xh.shape, dt), mg.pinned_empty(wh.shape, dt)
        px[...] = xh; pw[...] = wh
        pipe = mg.HostLayerPipeline(stride=1, padding=1, pool=(2, 2), pool_stride=2, chunks=chunks)
        out, dx, dw = pipe.step(px, pw)
        x, w = Tensor(xh), Tensor(wh)
        o = max_pool(conv_nd(x, w, stride=1, padding=1), (2, 2), 2)
        o.backward()
        np.testing.assert_array_equal(out, o.numpy())          # per-sample work: identical
        np.testing.assert_array_equal(dx, x.grad.get())
        assert_close(dw, w.grad.get(), dt)
        want_out, want_dx, want_dw = oc.layer_step(xh, wh, stride=1, padding=1, dilation=1,
                                                   pool=(2, 2), pool_stride=2)
        assert_close(out, want_out, dt); assert_close(dx, want_dx, dt); assert_close(dw, want_dw, dt)
        out2, _, _ = pipe.step(px, pw)                          # buffers are re-usable
        np.testing.assert_array_equal(out2, o.numpy())
        pipe.close()


@pytest.mark.parametrize("path", ["halo", "tma", "cpasync", "mma"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_kernels_do_not_write_outside_their_outputs(path, dt, monkeypatch):
    """compute-sanitizer is closed on this pool, so out-of-bounds WRITES are caught with guard
    bands: every output sits between two sentinel regions inside one big allocation."""
    monkeypatch.setenv("MGB_CONV_F32_PATH", "tma" if path == "halo" else path)
    monkeypatch.setenv("MGB_CONV_HALO", "1" if path == "halo" else "0")
    GUARD, SENT = 8192, -12345.0
    rng = np.random.default_rng(12)
    esz = np.dtype(dt).itemsize

    def guarded(n):
        big = mg.full((n + 2 * GUARD,), SENT, dt)
        return big, big.ptr + GUARD * esz

    def check_guard(big, n, what):
        h = big.get()
        assert np.all(h[:GUARD] == SENT) and np.all(h[GUARD + n:] == SENT), f"{what}: wrote outside its output"
        return h[GUARD:GUARD + n]

    for (N, Cc, F, X, W, s_, p_, d_) in [(3, 20, 36, (11, 13), (3, 3), 1, 1, 1), (2, 33, 17, (10, 11), (2, 3), 2, 1, 1),
                                         (2, 16, 130, (5, 6, 7), (2, 2, 2), 1, 0, 1), (5, 3, 4, (12,), (5,), 1, 2, 2)]:
        nd = len(X)
        tup = lambda v: (v,) * nd if isinstance(v, int) else tuple(v)
        xh = rng.standard_normal((N, Cc, *X)).astype(dt)
        wh = rng.standard_normal((F, Cc, *W)).astype(dt)
        G = oc.conv_out_shape(X, W, tup(s_), tup(p_), tup(d_))
        gh = rng.standard_normal((N, F, *G)).astype(dt)
        d = ConvDesc()
        d.dtype, d.nd, d.N, d.C, d.F = (0 if dt == np.float32 else 1), nd, N, Cc, F
        for i in range(nd):
            d.X[i], d.W[i], d.stride[i], d.pad[i], d.dil[i] = X[i], W[i], tup(s_)[i], tup(p_)[i], tup(d_)[i]
        xd, wd, gd = mg.asdevice(xh), mg.asdevice(wh), mg.asdevice(gh)
        for which, (a, b, want) in enumerate([
                (xd, wd, oc.conv_fwd(xh, wh, s_, p_, d_)),
                (gd, wd, oc.conv_dx(gh, xh.shape, wh, s_, p_, d_, dtype=dt)),
                (gd, xd, oc.conv_dw(gh, xh, wh.shape, s_, p_, d_))]):
            n = C.c_size_t(0)
            check(lib.mgb_conv_workspace_bytes(C.byref(d), which, C.byref(n)))
            ws_big, ws_ptr = guarded((n.value + 1024 + esz - 1) // esz + 64)   # room for the alignment shift
            ws_ptr = (ws_ptr + 1023) // 1024 * 1024          # workspaces are handed out 1 KB aligned
            out_big, out_ptr = guarded(want.size)
            fn = (lib.mgb_conv_fwd, lib.mgb_conv_dgrad, lib.mgb_conv_wgrad)[which]
            check(fn(C.byref(d), C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(out_ptr), C.c_void_p(ws_ptr), n, None))
            got = check_guard(out_big, want.size, f"conv[{which}] {path}")
            assert_close(got.reshape(want.shape), want, dt, err_msg=f"conv[{which}] {path}")
            h = ws_big.get()
            assert np.all(h[:GUARD] == SENT) and np.all(h[-GUARD:] == SENT), f"conv[{which}] {path}: workspace overrun"
    # pooling: vector, pair, generic and gather kernels
    for (shape, pool, stride) in [((3, 4, 16, 24), (2, 2), 2), ((2, 3, 10, 14), (2, 2), 2), ((2, 5, 9, 11), (3, 2), (2, 3)),
                                  ((2, 2, 4, 6, 8), (2, 2, 2), 2)]:
        xh = rng.standard_normal(shape).astype(dt)
        want = oc.max_pool_fwd(xh, pool, stride)
        gh = rng.standard_normal(want.shape).astype(dt)
        pd = PoolDesc()
        npool = len(pool)
        st = (stride,) * npool if isinstance(stride, int) else stride
        pd.dtype, pd.nd, pd.mode = (0 if dt == np.float32 else 1), npool, 0
        pd.lead = int(np.prod(shape[: len(shape) - npool]))
        for i in range(npool):
            pd.X[i], pd.P[i], pd.stride[i] = shape[len(shape) - npool + i], pool[i], st[i]
        xd, gd = mg.asdevice(xh), mg.asdevice(gh)
        y_big, y_ptr = guarded(want.size)
        check(lib.mgb_pool_fwd(C.byref(pd), C.c_void_p(xd.ptr), C.c_void_p(y_ptr), None))
        np.testing.assert_array_equal(check_guard(y_big, want.size, "pool fwd").reshape(want.shape), want)
        dx_big, dx_ptr = guarded(xh.size)
        check(lib.mgb_pool_bwd(C.byref(pd), C.c_void_p(gd.ptr), C.c_void_p(xd.ptr), C.c_void_p(dx_ptr), None))
        np.testing.assert_array_equal(check_guard(dx_big, xh.size, "pool bwd").reshape(shape),
                                      oc.max_pool_bwd(gh, xh, pool, stride).astype(dt))


def test_device_gradients_broadcast_and_zero_d_seeds():
    """`backward(grad)` broadcasts device gradients like host ones (tensor_base.py:1311-1330), and a 0-d
    terminal keeps a 0-d gradient"""
    rng = np.random.default_rng(4)
    xh, wh = rng.standard_normal((2, 2, 6, 6)), rng.standard_normal((3, 2, 3, 3))
    x, w = Tensor(xh), Tensor(wh)
    y = conv_nd(x, w, stride=1)
    y.backward(mg.asdevice(np.array(3.0)))                       # 0-d device gradient
    np.testing.assert_array_equal(y.grad.get(), np.full(y.shape, 3.0))
    y = conv_nd(x, w, stride=1)
    row = rng.standard_normal((y.shape[-1],))
    y.backward(mg.asdevice(row))                                  # trailing-axis device gradient
    np.testing.assert_array_equal(y.grad.get(), np.broadcast_to(row, y.shape))
    y = conv_nd(x, w, stride=1)
    with pytest.raises(ValueError):
        y.backward(mg.asdevice(np.ones((5,))))
    y.clear_graph()
    s = mg.einsum("ij,ij->", Tensor(xh[0, 0]), Tensor(xh[0, 1]))
    assert s.shape == ()
    s.backward(2.0)
    assert s.grad.shape == ()


def test_host_pipeline_validates_its_arguments():
    """raw pointers go to the copy engines: strided views are made contiguous, mixed dtypes size the
    outputs by the promoted dtype, and wrong out= / dx= / dw= buffers are rejected"""
    rng = np.random.default_rng(2)
    big = rng.standard_normal((6, 16, 12, 24)).astype(np.float32)
    xv = big[::2, :, :, ::2]                                     # non-contiguous view, shape (3, 16, 12, 12)
    w64 = rng.standard_normal((16, 16, 3, 3))                    # float64 filters with float32 x
    pipe = mg.HostLayerPipeline(stride=1, padding=1, pool=(2, 2), pool_stride=2, chunks=2)
    out, dx, dw = pipe.step(xv, w64)
    assert (out.dtype, dx.dtype, dw.dtype) == (np.float64, np.float32, np.float64)
    want = oc.layer_step(np.ascontiguousarray(xv), w64, stride=1, padding=1, dilation=1, pool=(2, 2), pool_stride=2)
    assert_close(out, want[0], np.float32); assert_close(dx, want[1], np.float32); assert_close(dw, want[2], np.float32)
    x32 = np.ascontiguousarray(xv)
    for bad in (dict(out=np.empty((3, 16, 6, 6), np.float32)),            # dtype of the promoted output is f64
                dict(dx=np.empty((3, 16, 12, 11), np.float32)),           # shape
                dict(dw=np.empty((16, 16, 3, 3, 2), np.float64)[..., 0]),  # not contiguous
                dict(out="nope")):
        with pytest.raises(ValueError):
            pipe.step(x32, w64, **bad)
    with pytest.raises(TypeError):
        pipe.step(x32.astype(np.int32), w64)
    with pytest.raises(ValueError):
        pipe.step(x32[:, :8], w64)                                # channel mismatch
    own = dict(out=np.empty((3, 16, 6, 6), np.float64), dx=np.empty(x32.shape, np.float32),
               dw=np.empty(w64.shape, np.float64))
    o2, dx2, dw2 = pipe.step(x32, w64, **own)
    assert o2 is own["out"] and dx2 is own["dx"] and dw2 is own["dw"]
    np.testing.assert_array_equal(o2, out)
    pipe.close()


def test_allocation_cache_respects_streams_and_devices():
    """a block that lived across a set_stream() is not recycled in allocation-stream order, and
    empty_cache() followed by new allocations still works"""
    import ctypes as C

    a = mg.empty((1 << 16,), np.float32)
    ptr_a = a.ptr
    del a
    b = mg.empty((1 << 16,), np.float32)
    assert b.ptr == ptr_a                                     # same stream, same size: recycled
    s = C.c_void_p()
    assert lib.mgb_stream_create(C.byref(s)) == 0
    try:
        mg.set_stream(s.value)
        b.fill(1.0)                                           # used on the other stream
        del b                                                 # freed behind an event, not cached for stream 0
        mg.set_stream(None)
        c = mg.empty((1 << 16,), np.float32)
        c.fill(2.0)
        assert float(c.get()[0]) == 2.0
    finally:
        mg.set_stream(None)
        lib.mgb_stream_destroy(s)
    mg.empty_cache()
    d = mg.full((1 << 16,), 3.0, np.float32)
    assert float(d.get()[-1]) == 3.0
