"""conv_pool: ``max_pool(conv_nd(x, w, ...), pool, pool_stride)`` as ONE Operation.

The reference materialises ``y = conv_nd(x, w)`` (src/mygrad/nnet/layers/conv.py:102-103), re-reads
it in ``MaxPoolND.__call__`` (nnet/layers/pooling.py:93-98) and a third time in
``MaxPoolND.backward_var`` (pooling.py:109-149); ``y`` and ``y.grad`` are the largest tensors of the
layer.  ``Tensor`` semantics expose both, so the drop-in pair ``conv_nd`` / ``max_pool`` has to write
them.  This opt-in Operation returns only the pooled tensor and keeps the rest on chip:

* forward — the tcgen05 convolution kernel pools its accumulator tile in the epilogue and writes the
  window maxima plus one byte per window, the position of its FIRST maximum in row-major window
  order (NaN counts as maximal: the reference's ``argmax`` rule, pooling.py:111);
* backward — ``(dp, argmax)`` become the channels-last hi/lo planes of ``dy`` directly
  (``mgb_pool_bwd_planes``), which dX and dW then read; ``dy`` never exists as an NCHW tensor.

Values, routing and both gradients are bit-identical to the unfused pair whenever ``conv_nd`` runs the
same (halo) convolution kernel — its choice at layer sizes like BASELINE config 2 — and agree to rounding
otherwise (tests/test_gpu_fused.py).
Geometries the fused kernels do not cover (float64, strided convolutions, pools other than 2x2/2,
fewer than 16 channels, 1-D / 3-D) run the same kernels as ``conv_nd`` + ``max_pool`` inside the op.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import numpy as np

from ..._lib import check, lib
from ...device_array import DeviceArray, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor
from .conv import ConvND, _ptr
from .pooling import MaxPoolND

__all__ = ["conv_pool", "ConvPoolND"]


class ConvPoolND(Operation):
    def __call__(self, x: Tensor, w: Tensor, *, stride, padding=0, dilation=1, pool, pool_stride) -> DeviceArray:
        self.variables = (x, w)
        conv, mp = ConvND(), MaxPoolND()
        y_shape = conv._plan(x, w, stride=stride, padding=padding, dilation=dilation)
        p_shape = mp._plan(y_shape, conv._dtype, pool, pool_stride)
        self._conv, self._pool = conv, mp
        # the reference's attributes of both ops
        self.padding, self.stride, self.dilation = conv.padding, conv.stride, conv.dilation
        self.pool, self.pool_stride = mp.pool, mp.stride
        self._dy_for: Optional[Tuple[int, object]] = None

        yes = C.c_int(0)
        if len(mp.pool):
            check(lib.mgb_conv_pool_supported(C.byref(conv._desc), C.byref(mp._desc), C.byref(yes)))
        self.fused = bool(yes.value)
        stream = C.c_void_p(current_stream())
        if self.fused:
            xin, win = conv._operands()
            conv._x_planes = conv._make_planes(0, xin)
            out = DeviceArray(p_shape, conv._dtype)
            self._argmax = DeviceArray(((out.size + 3) // 4,), np.float32)      # one byte per window
            ws, ws_bytes = conv._workspace(0)
            check(lib.mgb_conv_pool_fwd(C.byref(conv._desc), C.byref(mp._desc), C.c_void_p(xin.ptr),
                                        C.c_void_p(win.ptr), C.c_void_p(out.ptr), C.c_void_p(self._argmax.ptr),
                                        _ptr(conv._x_planes), C.c_void_p(ws.ptr if ws else 0),
                                        C.c_size_t(ws_bytes), stream))
            return out
        # not covered by the fused kernels: the same kernels as conv_nd + max_pool; y stays inside the op
        self._y = Tensor(conv(x, w, stride=stride, padding=padding, dilation=dilation), copy=False, constant=True)
        return mp(self._y, pool, pool_stride)

    def _dy(self, grad: DeviceArray):
        """dy of the convolution for this upstream gradient, made once per backward call:
        fused -> (None, channels-last planes); unfused -> (dy, None)"""
        if self._dy_for is not None and self._dy_for[0] == grad.ptr:
            return self._dy_for[1]
        conv, mp = self._conv, self._pool
        if self.fused:
            g = grad.astype(np.float32, copy=False)
            n = C.c_size_t(0)
            check(lib.mgb_conv_planes_bytes(C.byref(conv._desc), 1, C.byref(n)))
            planes = DeviceArray(((n.value + 7) // 8,), np.float64)
            check(lib.mgb_pool_bwd_planes(C.byref(conv._desc), C.byref(mp._desc), C.c_void_p(g.ptr),
                                          C.c_void_p(self._argmax.ptr), C.c_void_p(planes.ptr),
                                          C.c_void_p(current_stream())))
            res = (None, planes, g)     # g is kept alive until the kernel that reads it has been queued
        else:
            dy = mp.backward_var(grad, 0).astype(conv._dtype, copy=False)
            res = (dy, conv._make_planes(1, dy), None)
        self._dy_for = (grad.ptr, res)
        return res

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        dy, planes, _ = self._dy(grad)
        if index == 0:
            return self._conv._dgrad(dy, planes)
        return self._conv._wgrad(dy, planes)


def conv_pool(
    x,
    filter_bank,
    *,
    stride: Union[int, Tuple[int, ...]],
    padding: Union[int, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    pool: Tuple[int, ...],
    pool_stride: Union[int, Tuple[int, ...]],
    constant: Optional[bool] = None,
) -> Tensor:
    """``max_pool(conv_nd(x, filter_bank, stride=..., padding=..., dilation=...), pool, pool_stride)``
    with the convolution output kept on chip (see the module docstring).  Arguments, validation and
    exception types are those of ``conv_nd`` (conv.py:158-166, 373-395) and ``max_pool``
    (pooling.py:152-158)."""
    if x.ndim < 3:
        raise ValueError(f"`x` must possess at least three dimensions, got {x.ndim} dimensions")
    if x.ndim != filter_bank.ndim:
        raise ValueError(
            f"`x` ({x.ndim}-dimensions) must have the same dimensionality as "
            f"`filter_bank` ({filter_bank.ndim}-dimensions)"
        )
    if filter_bank.shape[1] != x.shape[1]:
        raise ValueError(
            f"`x.shape[1]` ({x.shape[1]}) must match `filter_bank.shape[1]` ({filter_bank.shape[1]})"
        )
    return Tensor._op(
        ConvPoolND,
        x,
        filter_bank,
        op_kwargs={"stride": stride, "padding": padding, "dilation": dilation, "pool": pool,
                   "pool_stride": pool_stride},
        constant=constant,
    )
