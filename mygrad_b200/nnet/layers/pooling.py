"""max_pool (and mean_pool) on B200: N-d window pooling over the trailing axes.

Drop-in for ``mygrad.nnet.layers.max_pool`` / ``MaxPoolND``
(src/mygrad/nnet/layers/pooling.py:12-149 op, :152-246 wrapper): same signature, same
attributes on the op (``variables, pool, stride``), AssertionError for malformed
pool/stride, ValueError for incompatible shapes, output ``(lead..., G0, ...)``.
The backward routes each window's gradient to the FIRST maximal element in row-major
window order (NaN counts as maximal), accumulating overlaps in float64 in increasing
window order — bit-identical to the reference's argmax + np.add.at.

``mean_pool`` has no counterpart in the reference: its oracle is the
composition ``sliding_window_view(x, pool, stride).mean(axis=pool_axes)``.
"""
from __future__ import annotations

import ctypes as C
import os
from numbers import Integral
from typing import Optional, Tuple, Union

import numpy as np

from ..._lib import PoolDesc, check, lib
from ...device_array import DeviceArray, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor

__all__ = ["max_pool", "mean_pool", "MaxPoolND", "MeanPoolND"]

_MAX_SPATIAL = 3


class MaxPoolND(Operation):
    _mode = 0  # 0 = max, 1 = mean

    def __call__(self, x: Tensor, pool, stride) -> DeviceArray:
        self.variables = (x,)
        xd = x.data
        out_shape = self._plan(xd.shape, xd.dtype, pool, stride)
        if len(self.pool) == 0:  # pooling over no axes is the identity
            return xd.copy()
        y = DeviceArray(out_shape, xd.dtype)
        check(lib.mgb_pool_fwd(C.byref(self._desc), C.c_void_p(xd.ptr), C.c_void_p(y.ptr),
                               C.c_void_p(current_stream())))
        return y

    def _plan(self, in_shape, dtype, pool, stride) -> Tuple[int, ...]:
        """validation, the op's public attributes and the C descriptor; returns the output shape"""

        assert isinstance(pool, (tuple, list, np.ndarray)) and all(
            isinstance(i, Integral) and i >= 0 for i in pool
        )
        pool = tuple(int(i) for i in pool)
        assert all(i > 0 for i in pool)
        assert len(in_shape) >= len(pool), (
            "The number of pooled dimensions cannot exceed the dimensionality of the data."
        )
        num_pool = len(pool)
        if isinstance(stride, Integral):
            stride = (int(stride),) * num_pool
        else:
            stride = tuple(stride)
        assert len(stride) == num_pool and all(isinstance(s, Integral) and s >= 1 for s in stride)
        stride = tuple(int(s) for s in stride)

        self.pool = np.asarray(pool, dtype=int)
        self.stride = np.asarray(stride, dtype=int)

        lead_shape = tuple(in_shape[: len(in_shape) - num_pool])
        x_shape = tuple(in_shape[len(in_shape) - num_pool:])
        out_shape = []
        ok = True
        for X, P, s in zip(x_shape, pool, stride):
            num = X - P
            ok = ok and num >= 0 and num % s == 0
            out_shape.append(num // s + 1)
        if not ok:
            raise ValueError(
                "Stride and kernel dimensions are incompatible: \n"
                f"Input dimensions: {tuple(x_shape)}\n"
                f"Stride dimensions: {tuple(stride)}\n"
                f"Pooling dimensions: {tuple(pool)}\n"
            )
        if num_pool > _MAX_SPATIAL:
            raise NotImplementedError(
                f"the B200 path pools over at most {_MAX_SPATIAL} trailing axes, got {num_pool}"
            )

        d = PoolDesc()
        d.dtype = 0 if np.dtype(dtype) == np.float32 else 1
        d.nd = num_pool
        d.mode = self._mode
        lead = 1
        for s in lead_shape:
            lead *= s
        d.lead = lead
        for i in range(num_pool):
            d.X[i], d.P[i], d.stride[i] = x_shape[i], pool[i], stride[i]
        self._desc = d
        return (*lead_shape, *out_shape)

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        xd = self.variables[index].data
        if len(self.pool) == 0:
            return grad.copy()
        g = grad.astype(xd.dtype, copy=False)
        dx = DeviceArray(xd.shape, xd.dtype)
        if self._feeds_conv_planes(g, xd, dx):
            return dx
        check(lib.mgb_pool_bwd(C.byref(self._desc), C.c_void_p(g.ptr), C.c_void_p(xd.ptr),
                               C.c_void_p(dx.ptr), C.c_void_p(current_stream())))
        return dx


    def _feeds_conv_planes(self, g: DeviceArray, yd: DeviceArray, dy: DeviceArray) -> bool:
        """The transparent half of the conv -> pool fusion for the drop-in pair
        ``max_pool(conv_nd(...))``: when the pooled tensor was produced by a float32 tensor-core convolution and
        nothing else consumes it, one kernel writes dy — the public ``y.grad``, bit-identical to ``mgb_pool_bwd``'s —
        AND the channels-last hi/lo planes of dy that the convolution's dX / dW read, instead of a second pass
        over dy (``mgb_conv_make_planes``).  The planes are handed to the ConvND op keyed by dy's pointer."""
        from .conv import ConvND

        y = self.variables[0]
        conv = y.creator
        if self._mode != 0 or not isinstance(conv, ConvND) or len(y._ops) != 1 or yd.dtype != np.float32:
            return False
        if os.environ.get("MGB_POOL_FEEDS_CONV", "1") == "0":      # A/B switch for tests and measurements
            return False
        yes = C.c_int(0)
        check(lib.mgb_conv_pool_supported(C.byref(conv._desc), C.byref(self._desc), C.byref(yes)))
        if not yes.value:
            return False
        n = C.c_size_t(0)
        check(lib.mgb_conv_planes_bytes(C.byref(conv._desc), 1, C.byref(n)))
        planes = DeviceArray(((n.value + 7) // 8,), np.float64)
        check(lib.mgb_pool_bwd_dy_planes(C.byref(conv._desc), C.byref(self._desc), C.c_void_p(g.ptr),
                                         C.c_void_p(yd.ptr), C.c_void_p(dy.ptr), C.c_void_p(planes.ptr),
                                         C.c_void_p(current_stream())))
        conv._dy_planes = (dy.ptr, planes)
        return True


class MeanPoolND(MaxPoolND):
    _mode = 1


def max_pool(
    x,
    pool: Tuple[int, ...],
    stride: Union[int, Tuple[int, ...]],
    *,
    constant: Optional[bool] = None,
) -> Tensor:
    """Max-pooling over the last ``len(pool)`` axes of `x` with the given window and stride;
    only 'valid' window placements (pooling.py:152-246)."""
    return Tensor._op(MaxPoolND, x, op_args=(pool, stride), constant=constant)


def mean_pool(
    x,
    pool: Tuple[int, ...],
    stride: Union[int, Tuple[int, ...]],
    *,
    constant: Optional[bool] = None,
) -> Tensor:
    """Mean-pooling with the same window semantics as :func:`max_pool`."""
    return Tensor._op(MeanPoolND, x, op_args=(pool, stride), constant=constant)
