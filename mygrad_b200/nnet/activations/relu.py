"""relu on the device — drop-in for ``mygrad.nnet.activations.relu``
(src/mygrad/nnet/activations/relu.py:10-17): ``y = x * (x > 0)`` (so negatives give -0.0 and
NaN stays NaN), ``dx = grad * (x > 0)``; the mask is recomputed from x in the backward."""
from __future__ import annotations

import ctypes as C
from typing import Optional

from ..._lib import check, lib
from ...device_array import DeviceArray, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor

__all__ = ["relu", "ReLu"]


class ReLu(Operation):
    def __call__(self, a: Tensor) -> DeviceArray:
        self.variables = (a,)
        x = a.data
        y = DeviceArray(x.shape, x.dtype)
        check(lib.mgb_relu_fwd(x.code, C.c_size_t(x.size), C.c_void_p(x.ptr), C.c_void_p(y.ptr),
                               C.c_void_p(current_stream())))
        return y

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        x = self.variables[0].data
        g = grad.astype(x.dtype, copy=False)
        dx = DeviceArray(x.shape, x.dtype)
        check(lib.mgb_relu_bwd(x.code, C.c_size_t(x.size), C.c_void_p(g.ptr), C.c_void_p(x.ptr),
                               C.c_void_p(dx.ptr), C.c_void_p(current_stream())))
        return dx


def relu(x, *, constant: Optional[bool] = None) -> Tensor:
    """rectified linear unit, element-wise: ``x if x > 0 else 0``"""
    return Tensor._op(ReLu, x, constant=constant)
