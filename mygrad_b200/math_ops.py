"""The three graph ops a conv net needs around conv_nd / max_pool:

* ``add``     — ``mygrad.add`` (Add, src/mygrad/math/arithmetic/ops.py:28-32): NumPy broadcasting;
  backward hands ``grad`` itself to ``Operation.backward``, which copies and un-broadcasts.
* ``matmul``  — ``mygrad.matmul`` for 2-D operands (MatMul, src/mygrad/math/misc/ops.py:91-127),
  executed as a 1x1 convolution on the conv GEMM kernels: ``A (N,D) @ B (D,M)`` is
  ``conv(x=(N,D,1), w=B^T (M,D,1))``; dA is the conv dX, dB the transposed conv dW.
* ``reshape`` — ``mygrad.reshape`` (Reshape, src/mygrad/tensor_manip/array_shape/ops.py:50-66):
  metadata-only view; backward reshapes the gradient back.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import ConvDesc, check, lib
from .device_array import DeviceArray, current_stream
from .nnet.layers.conv import _scratch
from .operation_base import Operation
from .tensor_base import Tensor

__all__ = ["add", "matmul", "reshape", "Add", "MatMul", "Reshape"]


def _i64(values):
    values = tuple(int(v) for v in values)
    return (C.c_int64 * max(1, len(values)))(*values)


class Add(Operation):
    def __call__(self, a: Tensor, b: Tensor) -> DeviceArray:
        self.variables = (a, b)
        x, y = a.data, b.data
        dtype = np.result_type(x.dtype, y.dtype)
        x, y = x.astype(dtype, copy=False), y.astype(dtype, copy=False)
        shape = np.broadcast_shapes(x.shape, y.shape)   # raises ValueError like NumPy
        nd = len(shape)
        xs = (1,) * (nd - x.ndim) + x.shape
        ys = (1,) * (nd - y.ndim) + y.shape
        out = DeviceArray(shape, dtype)
        check(lib.mgb_add_broadcast(out.code, nd, _i64(shape), _i64(xs), _i64(ys), C.c_void_p(x.ptr),
                                    C.c_void_p(y.ptr), C.c_void_p(out.ptr), C.c_void_p(current_stream())))
        return out

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        return grad


class Reshape(Operation):
    can_return_view = True

    def __call__(self, a: Tensor, shape) -> DeviceArray:
        self.variables = (a,)
        return a.data.reshape(shape)

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        return grad.reshape(self.variables[0].shape)


class MatMul(Operation):
    """2-D ``a @ b`` as a 1x1 convolution (any dtype mix promotes like NumPy)."""

    def __call__(self, a: Tensor, b: Tensor) -> DeviceArray:
        self.variables = (a, b)
        x, y = a.data, b.data
        if x.ndim != 2 or y.ndim != 2:
            raise NotImplementedError("the device matmul handles 2-D @ 2-D operands")
        if x.shape[1] != y.shape[0]:
            raise ValueError(f"matmul: shapes {x.shape} and {y.shape} are not aligned")
        self._dtype = np.result_type(x.dtype, y.dtype)
        n, d = x.shape
        m = y.shape[1]
        self._skinny = m <= 32
        if self._skinny:   # classifier-head shape: bandwidth-bound kernels instead of GEMM tiles
            self._a = x.astype(self._dtype, copy=False)
            self._b = y.astype(self._dtype, copy=False)
            out = DeviceArray((n, m), self._dtype)
            check(lib.mgb_matmul_skinny(out.code, 0, n, d, m, C.c_void_p(self._a.ptr), C.c_void_p(self._b.ptr),
                                        C.c_void_p(out.ptr), C.c_void_p(current_stream())))
            return out
        desc = self._desc = ConvDesc()
        desc.dtype, desc.nd = (0 if self._dtype == np.float32 else 1), 1
        desc.N, desc.C, desc.F = n, d, m
        desc.X[0], desc.W[0], desc.stride[0], desc.pad[0], desc.dil[0] = 1, 1, 1, 0, 1
        self._a = x.astype(self._dtype, copy=False)
        self._bt = self._transpose(y.astype(self._dtype, copy=False))          # (M, D) filter bank
        out = DeviceArray((n, m), self._dtype)
        ws, nbytes = self._ws(0)
        check(lib.mgb_conv_fwd(C.byref(desc), C.c_void_p(self._a.ptr), C.c_void_p(self._bt.ptr),
                               C.c_void_p(out.ptr), C.c_void_p(ws.ptr if ws else 0), C.c_size_t(nbytes),
                               C.c_void_p(current_stream())))
        return out

    @staticmethod
    def _transpose(src: DeviceArray) -> DeviceArray:
        rows, cols = src.shape
        dst = DeviceArray((cols, rows), src.dtype)
        check(lib.mgb_transpose2d(src.code, rows, cols, C.c_void_p(src.ptr), C.c_void_p(dst.ptr),
                                  C.c_void_p(current_stream())))
        return dst

    def _ws(self, which: int):
        n = C.c_size_t(0)
        check(lib.mgb_conv_workspace_bytes(C.byref(self._desc), which, C.byref(n)))
        return (_scratch(n.value), n.value) if n.value else (None, 0)

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        a, b = (v.data for v in self.variables)
        g = grad.astype(self._dtype, copy=False)
        stream = C.c_void_p(current_stream())
        if self._skinny:
            n, d = a.shape
            m = b.shape[1]
            out = DeviceArray(a.shape if index == 0 else b.shape, self._dtype)
            lhs, rhs = (g, self._b) if index == 0 else (self._a, g)
            check(lib.mgb_matmul_skinny(out.code, 1 if index == 0 else 2, n, d, m, C.c_void_p(lhs.ptr),
                                        C.c_void_p(rhs.ptr), C.c_void_p(out.ptr), stream))
            return out
        if index == 0:      # dA = grad @ B^T  == conv dX
            da = DeviceArray(a.shape, self._dtype)
            ws, nbytes = self._ws(1)
            check(lib.mgb_conv_dgrad(C.byref(self._desc), C.c_void_p(g.ptr), C.c_void_p(self._bt.ptr),
                                     C.c_void_p(da.ptr), C.c_void_p(ws.ptr if ws else 0), C.c_size_t(nbytes),
                                     stream))
            return da
        dbt = DeviceArray((b.shape[1], b.shape[0]), self._dtype)   # conv dW = (A^T grad)^T
        ws, nbytes = self._ws(2)
        check(lib.mgb_conv_wgrad(C.byref(self._desc), C.c_void_p(g.ptr), C.c_void_p(self._a.ptr),
                                 C.c_void_p(dbt.ptr), C.c_void_p(ws.ptr if ws else 0), C.c_size_t(nbytes),
                                 stream))
        return self._transpose(dbt)


def add(a, b, *, constant: Optional[bool] = None) -> Tensor:
    return Tensor._op(Add, a, b, constant=constant)


def matmul(a, b, *, constant: Optional[bool] = None) -> Tensor:
    return Tensor._op(MatMul, a, b, constant=constant)


def reshape(a, *newshape, constant: Optional[bool] = None) -> Tensor:
    if len(newshape) == 1 and isinstance(newshape[0], (tuple, list)):
        newshape = tuple(newshape[0])
    return Tensor._op(Reshape, a, op_args=(newshape,), constant=constant)
