"""Host-side logic that needs no GPU: argument validation with the reference's exception
types, graph bookkeeping of Tensor._op / backward, shard arithmetic."""
import numpy as np
import pytest

import mygrad_b200 as mg
from mygrad_b200 import DeviceArray, Operation, Tensor
from mygrad_b200.dist import scale_for_mean_loss, shard_bounds
from mygrad_b200.nnet.layers import conv_nd, max_pool


def meta(shape, dtype=np.float64, constant=None) -> Tensor:
    """shape-only tensor (null device pointer): enough to drive validation code"""
    return Tensor(DeviceArray(shape, dtype, _ptr=0), copy=False, constant=constant)


def test_conv_wrapper_value_errors():
    # conv.py:373-387 / tests/nnet/layers/test_conv.py:21-55, 220-228
    with pytest.raises(ValueError):
        conv_nd(np.zeros((2, 2)), np.zeros((1, 2, 2)), stride=1)           # ndim < 3
    with pytest.raises(ValueError):
        conv_nd(np.zeros((1, 2, 2, 2)), np.zeros((1, 2, 2)), stride=1)     # ndim mismatch
    with pytest.raises(ValueError):
        conv_nd(np.zeros((1, 2, 2, 2)), np.zeros((1, 3, 2, 2)), stride=1)  # channel mismatch


def test_conv_op_assertion_and_value_errors():
    # tests/nnet/layers/test_conv.py:231-244
    x, w = meta((1, 2, 2, 2)), meta((1, 2, 2, 2))
    with pytest.raises(ValueError):
        conv_nd(x, meta((1, 2, 3, 2)), stride=1, padding=0)   # filter larger than input
    for kwargs in (dict(stride=0), dict(stride=[1, 2, 3]), dict(stride=1, padding=-1),
                   dict(stride=1, padding=[1, 2, 3]), dict(stride=1, dilation=0),
                   dict(stride=1, dilation=(1, 1, 1))):
        with pytest.raises(AssertionError):
            conv_nd(x, w, **kwargs)
    with pytest.raises(ValueError) as e:
        conv_nd(x, w, stride=3, padding=1)                    # non-integer output extent
    assert "Stride and kernel dimensions are incompatible" in str(e.value)
    with pytest.raises(ValueError):   # BASELINE.json config 3 as written
        conv_nd(meta((1, 2, 64, 64)), meta((3, 2, 5, 5)), stride=2, dilation=2)


def test_pool_assertion_and_value_errors():
    # tests/nnet/layers/test_maxpool.py:181-193
    x = meta((1, 2, 2, 2), np.float32)
    with pytest.raises(ValueError):
        max_pool(x, (3,) * 3, (1,) * 3)
    with pytest.raises(AssertionError):
        max_pool(x, (2,) * 3, (0,) * 3)
    with pytest.raises(AssertionError):
        max_pool(x, (2,) * 2, [1, 2, 3])
    with pytest.raises(ValueError):
        max_pool(x, (1,) * 3, (3,) * 3)
    with pytest.raises(AssertionError):
        max_pool(x, (0, 2), 1)
    with pytest.raises(AssertionError):
        max_pool(x, 2, 1)                                     # pool must be a sequence
    with pytest.raises(AssertionError):
        max_pool(x, (1,) * 5, 1)                              # more pooled axes than dims


class _Identity(Operation):
    """backward_var returns `grad` itself: forces the defensive copy in Operation.backward"""

    def __call__(self, a):
        self.variables = (a,)
        return a.data

    def backward_var(self, grad, index, **kwargs):
        return grad


def test_op_graph_bookkeeping_without_device():
    a = meta((2, 3))
    a._grad = DeviceArray((2, 3), np.float64, _ptr=0)
    out = Tensor._op(_Identity, a)
    assert a.grad is None                       # ops null their inputs' grads (tensor_base.py:1163-1172)
    assert out.creator is not None and out.creator.variables == (a,)
    assert len(a._ops) == 1 and out.constant is False
    assert Tensor._op(_Identity, a, constant=True).constant is True
    assert Tensor._op(_Identity, meta((2,), constant=True)).constant is True   # all-constant inputs
    out.clear_graph()
    assert out.creator is None and not a._ops
    with mg.no_autodiff:
        o2 = Tensor._op(_Identity, a)
    assert o2.creator is None and not a._ops    # no graph tracking inside no_autodiff
    assert mg._graph_tracking.TRACK_GRAPH is True

    @mg.no_autodiff
    def f():
        return mg._graph_tracking.TRACK_GRAPH

    assert f() is False and mg._graph_tracking.TRACK_GRAPH is True


def test_tensor_constructor_rules():
    with pytest.raises(TypeError):
        Tensor(DeviceArray((2,), np.float64, _ptr=0), constant="yes", copy=False)
    with pytest.raises(TypeError):
        DeviceArray((2,), np.int32, _ptr=0)
    t = meta((3, 4, 5), np.float32)
    assert (t.shape, t.ndim, t.size, t.dtype, t.constant) == ((3, 4, 5), 3, 60, np.float32, False)
    v = t.data.reshape(12, -1)
    assert v.shape == (12, 5) and v.base is t.data and v.ptr == t.data.ptr
    with pytest.raises(ValueError):
        t.data.reshape(7, -1)


def test_device_array_interface_dict():
    d = DeviceArray((2, 3), np.float32, _ptr=4096)
    cai = d.__cuda_array_interface__
    assert cai["shape"] == (2, 3) and cai["typestr"] == "<f4" and cai["data"] == (4096, False)
    assert cai["version"] == 3 and cai["strides"] is None


def test_shard_bounds_cover_the_batch():
    for n in (0, 1, 7, 32, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_bounds(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(32, 3, 8) == (12, 16)                 # config 4: 4 samples per GPU
    with pytest.raises(ValueError):
        shard_bounds(8, 8, 8)
    assert scale_for_mean_loss(256, 2048) == 0.125
