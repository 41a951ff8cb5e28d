"""BASELINE.json config 5: a small CNN (2x[conv_nd + bias + relu + max_pool] + dense + softmax
cross-entropy) through Tensor.backward, entirely on the device, against golden vectors from the
live reference (tests/golden/make_golden.py::run_cnn_case) — and the ops around the path one
by one against NumPy restatements of the reference formulas."""
import numpy as np
import pytest

import mygrad_b200 as mg
from conftest import GOLDEN, assert_close
from mygrad_b200 import Tensor
from mygrad_b200.nnet import conv_nd, max_pool, relu, softmax_crossentropy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_small_cnn_vs_reference_golden(tag):
    g = np.load(GOLDEN / "cnn_case.npz")
    c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(tag + "/")}
    dt = c["x"].dtype
    p = {k: Tensor(c[k]) for k in ("w1", "b1", "w2", "b2", "wd", "bd")}
    x = Tensor(c["x"])
    h = max_pool(relu(conv_nd(x, p["w1"], stride=1, padding=1) + p["b1"]), (2, 2), 2)
    h = max_pool(relu(conv_nd(h, p["w2"], stride=1, padding=1) + p["b2"]), (2, 2), 2)
    logits = h.reshape(8, -1) @ p["wd"] + p["bd"]
    loss = softmax_crossentropy(logits, c["labels"])
    assert loss.shape == () and loss.dtype == dt
    loss.backward()
    assert_close(logits.numpy(), c["logits"], dt)
    assert_close(loss.numpy(), c["loss"], dt)
    for k in p:
        assert p[k].grad.shape == c[k].shape and p[k].grad.dtype == dt
        assert_close(p[k].grad.get(), c["d" + k], dt, err_msg=f"{tag} d{k}")
    assert_close(x.grad.get(), c["dx"], dt, err_msg=f"{tag} dx")
    assert loss.creator is None            # graph cleared


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_relu_semantics(dt):
    xh = np.array([-2.0, -0.0, 0.0, 3.5, np.nan, np.inf, -np.inf, 1e-30], dtype=dt)
    x = Tensor(xh)
    y = relu(x)
    with np.errstate(invalid="ignore"):
        want = xh * (xh > 0).astype(dt)        # src/mygrad/nnet/activations/relu.py:10-14
        np.testing.assert_array_equal(y.numpy(), want)
        assert np.signbit(y.numpy()[0]) and np.signbit(want[0])        # -2 * 0 = -0.0, like the reference
        gh = np.arange(1, 9, dtype=dt)
        y.backward(gh)
        np.testing.assert_array_equal(x.grad.get(), gh * (xh > 0).astype(dt))
    big = np.random.default_rng(0).standard_normal((7, 1031)).astype(dt)   # odd size: vector body + tail
    np.testing.assert_array_equal(relu(Tensor(big)).numpy(), big * (big > 0))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_add_broadcasting_and_unbroadcast(dt):
    rng = np.random.default_rng(1)
    for sa, sb in [((4, 5, 6, 7), (1, 5, 1, 1)), ((4, 5, 6, 7), (5, 1, 1)), ((3, 8), (8,)), ((3, 8), (3, 1)),
                   ((2, 3, 4), (2, 3, 4)), ((5, 1, 4), (1, 6, 1)), ((7,), ()), ((1, 3, 1, 2), (4, 1, 5, 1))]:
        ah, bh = rng.standard_normal(sa).astype(dt), rng.standard_normal(sb).astype(dt)
        a, b = Tensor(ah), Tensor(bh)
        out = a + b
        np.testing.assert_array_equal(out.numpy(), ah + bh)
        gh = rng.standard_normal(out.shape).astype(dt)
        out.backward(gh)
        from oracle import conv_pool_oracle as oc
        assert_close(a.grad.get(), oc.reduce_broadcast(gh, sa), dt)
        assert_close(b.grad.get(), np.asarray(oc.reduce_broadcast(gh, sb)), dt)
    with pytest.raises(ValueError):
        Tensor(np.ones((2, 3), dt)) + Tensor(np.ones((4,), dt))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_matmul_2d(dt):
    rng = np.random.default_rng(2)
    for n, d, m in [(5, 7, 3), (64, 96, 10), (33, 130, 48), (2048 // 8, 4096 // 8, 10)]:
        ah, bh = rng.standard_normal((n, d)).astype(dt), rng.standard_normal((d, m)).astype(dt)
        a, b = Tensor(ah), Tensor(bh)
        out = a @ b
        a64, b64 = ah.astype(np.float64), bh.astype(np.float64)
        assert_close(out.numpy(), a64 @ b64, dt)
        gh = rng.standard_normal((n, m)).astype(dt)
        out.backward(gh)
        assert_close(a.grad.get(), gh.astype(np.float64) @ b64.T, dt)      # math/misc/ops.py:107-113
        assert_close(b.grad.get(), a64.T @ gh.astype(np.float64), dt)      # math/misc/ops.py:115-122
    with pytest.raises(ValueError):
        Tensor(np.ones((2, 3), dt)) @ Tensor(np.ones((4, 2), dt))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_softmax_crossentropy(dt):
    rng = np.random.default_rng(3)
    for n, c in [(4, 3), (64, 10), (37, 100)]:
        xh = (rng.standard_normal((n, c)) * 3).astype(dt)
        labels = rng.integers(0, c, n)
        x = Tensor(xh)
        loss = softmax_crossentropy(x, labels)
        x64 = xh.astype(np.float64)
        lse = np.log(np.exp(x64 - x64.max(1, keepdims=True)).sum(1, keepdims=True)) + x64.max(1, keepdims=True)
        logp = x64 - lse
        assert_close(loss.numpy(), np.asarray(-logp[np.arange(n), labels].mean()), dt)
        loss.backward()
        back = np.exp(logp)
        back[np.arange(n), labels] -= 1.0
        assert_close(x.grad.get(), back / n, dt)
    with pytest.raises(ValueError):
        softmax_crossentropy(Tensor(np.ones((3, 4), dt)), [0, 1])
    # label conventions of the reference: fancy indexing x[range(N), y] (softmax_crossentropy.py:41) accepts
    # negative indices and raises IndexError out of range; check_loss_inputs raises TypeError for float labels
    with pytest.raises(IndexError):
        softmax_crossentropy(Tensor(np.ones((3, 4), dt)), [0, 1, 4])
    with pytest.raises(TypeError):
        softmax_crossentropy(Tensor(np.ones((3, 4), dt)), [0.0, 1.0, 2.0])
    xs = np.random.default_rng(3).standard_normal((3, 4)).astype(dt)
    a = softmax_crossentropy(Tensor(xs), [-1, 0, -4])
    b = softmax_crossentropy(Tensor(xs), [3, 0, 0])
    assert a.item() == b.item()


def test_reshape_is_a_view_and_backprops():
    xh = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    x = Tensor(xh)
    r = x.reshape(6, -1)
    assert r.shape == (6, 4) and r.data.ptr == x.data.ptr
    r.backward(np.arange(24.0).reshape(6, 4))
    np.testing.assert_array_equal(x.grad.get(), xh)


def test_view_ops_do_not_null_the_grads_of_their_inputs():
    """tensor_base.py:1134-1172: only NON-view ops clear the grads of their inputs"""
    from mygrad_b200.nnet.layers import conv_nd

    rng = np.random.default_rng(1)
    x = Tensor(rng.standard_normal((2, 3, 4)))
    (x + 1.0).backward()
    assert x.grad is not None
    _ = x.reshape(6, 4)                 # a view: x.grad survives
    assert x.grad is not None
    _ = x + 2.0                         # not a view: nulled
    assert x.grad is None
