"""batchnorm on B200: normalisation over every axis but axis 1, with optional affine map.

Drop-in for ``mygrad.nnet.layers.batchnorm`` / ``BatchNorm``
(src/mygrad/nnet/layers/batchnorm.py:13-111 op, :114-175 wrapper): same signature
(``gamma`` / ``beta`` optional, ``eps`` keyword-only and required), same attributes on the op
(``variables, gamma, beta, mean, var``), same gradients:

    dx     = gamma/std * (g - mean(g) - x_norm * mean(g * x_norm))      (batchnorm.py:76-99)
    dgamma = sum(g * x_norm)                                            (batchnorm.py:101-102)
    dbeta  = sum(g)                                                     (batchnorm.py:104-107)

``x_norm`` is not kept (the reference stores it, batchnorm.py:63): the backward recomputes it from
x and the saved float64 moments, which costs one extra read of x instead of a write + read of a
tensor the size of x.  With ``dp=DataParallel`` the moments and the two backward sums are merged
across ranks (NCCL), so that every rank normalises with the statistics of the FULL batch — the
result to match is the single-process reference on the concatenated batch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ..._lib import check, lib
from ...device_array import DeviceArray, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor

__all__ = ["batchnorm", "BatchNorm"]


def _vp(a: Optional[DeviceArray]) -> C.c_void_p:
    return C.c_void_p(a.ptr if a is not None else 0)


class BatchNorm(Operation):
    """``mean`` and ``var`` (shape (C,), x's dtype) are bound when the op is called, as in the reference."""

    def __call__(self, x: Tensor, gamma: Tensor, beta: Tensor, *, eps, dp=None) -> DeviceArray:
        self.variables = (x, gamma, beta)
        xd = x.data
        if xd.ndim < 2:
            raise ValueError(f"`x` must have at least two dimensions (N, C, ...), got shape {xd.shape}")
        self.gamma = gamma if gamma.size else None     # empty array == "not supplied" (batchnorm.py:47-50)
        self.beta = beta if beta.size else None
        n, c = xd.shape[0], xd.shape[1]
        s = int(np.prod(xd.shape[2:], dtype=np.int64)) if xd.ndim > 2 else 1
        for name, p in (("gamma", self.gamma), ("beta", self.beta)):
            if p is not None and p.size != c:
                raise ValueError(f"`{name}` must have shape ({c},), got {p.shape}")
        if n * s == 0 and c:
            raise ValueError("batchnorm needs at least one element per channel")
        self._dims = (n, c, s)
        self._eps = float(eps)
        self._dp = dp if dp is not None and dp.world_size > 1 else None
        dt = xd.dtype
        stream = C.c_void_p(current_stream())

        nbytes = C.c_size_t(0)
        check(lib.mgb_batchnorm_workspace_bytes(n, c, C.byref(nbytes)))
        self._ws = DeviceArray(((nbytes.value + 7) // 8,), np.float64)
        self._moments = DeviceArray((3, c), np.float64)     # mean, population variance, count
        check(lib.mgb_batchnorm_stats(xd.code, n, c, s, _vp(xd), _vp(self._moments), _vp(self._ws),
                                      nbytes, stream))
        if self._dp is not None:
            self._merge_moments()
        g = self.gamma.data.astype(dt, copy=False) if self.gamma is not None else None
        b = self.beta.data.astype(dt, copy=False) if self.beta is not None else None
        y = DeviceArray(xd.shape, dt)
        check(lib.mgb_batchnorm_fwd(xd.code, n, c, s, _vp(xd), _vp(self._moments), _vp(g), _vp(b),
                                    C.c_double(self._eps), _vp(y), stream))
        self._sums = None
        self._sums_for = None
        return y

    # -- the reference's attributes ----------------------------------------------------------
    def _moment(self, row: int) -> DeviceArray:
        c = self._dims[1]
        view = DeviceArray((c,), np.float64, _ptr=self._moments.ptr + row * c * 8, _base=self._moments)
        return view.astype(self.variables[0].data.dtype)

    @property
    def mean(self) -> DeviceArray:
        return self._moment(0)

    @property
    def var(self) -> DeviceArray:
        return self._moment(1)

    # -- data parallel -------------------------------------------------------------------------
    def _merge_moments(self) -> None:
        c = self._dims[1]
        stream = C.c_void_p(current_stream())
        buf = DeviceArray((3 * c,), np.float64)
        head = DeviceArray((2 * c,), np.float64, _ptr=buf.ptr, _base=buf)
        tail = DeviceArray((c,), np.float64, _ptr=buf.ptr + 2 * c * 8, _base=buf)
        check(lib.mgb_batchnorm_merge(0, c, _vp(self._moments), _vp(buf), stream))
        self._dp.allreduce([head])
        check(lib.mgb_batchnorm_merge(1, c, _vp(self._moments), _vp(buf), stream))
        self._dp.allreduce([tail])
        check(lib.mgb_batchnorm_merge(2, c, _vp(self._moments), _vp(buf), stream))

    # -- backward ----------------------------------------------------------------------------------
    def _reduce(self, grad: DeviceArray) -> DeviceArray:
        """sum(g) and sum(g * x_norm) per channel, computed once per incoming gradient"""
        if self._sums is not None and self._sums_for == grad.ptr:
            return self._sums
        n, c, s = self._dims
        xd = self.variables[0].data
        sums = DeviceArray((2, c), np.float64)
        check(lib.mgb_batchnorm_bwd_reduce(xd.code, n, c, s, _vp(grad), _vp(xd), _vp(self._moments),
                                           C.c_double(self._eps), _vp(sums), _vp(self._ws),
                                           C.c_size_t(self._ws.nbytes), C.c_void_p(current_stream())))
        if self._dp is not None:
            self._dp.allreduce([sums])
        self._sums, self._sums_for = sums, grad.ptr
        return sums

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        n, c, s = self._dims
        xd = self.variables[0].data
        dt = xd.dtype
        g = grad.astype(dt, copy=False)
        sums = self._reduce(g)
        if index == 0:
            gm = self.gamma.data.astype(dt, copy=False) if self.gamma is not None else None
            dx = DeviceArray(xd.shape, dt)
            check(lib.mgb_batchnorm_bwd_dx(xd.code, n, c, s, _vp(g), _vp(xd), _vp(self._moments), _vp(sums),
                                           _vp(gm), C.c_double(self._eps), _vp(dx),
                                           C.c_void_p(current_stream())))
            return dx
        row = 1 if (index == 1 and self.gamma is not None) else 0
        if not ((index == 1) or index == 2):
            raise IndexError(index)
        view = DeviceArray((c,), np.float64, _ptr=sums.ptr + row * c * 8, _base=sums)
        out = view.astype(dt)
        # with dp= the channel sums were all-reduced above: dgamma / dbeta already cover the global batch,
        # and DataParallel.allreduce must not sum them again
        out.reduced = self._dp is not None and self._dp.world_size > 1
        return out


def batchnorm(x, *, gamma=None, beta=None, eps: float, constant: Optional[bool] = None, dp=None) -> Tensor:
    """``gamma * (x - E[x]) / sqrt(Var[x] + eps) + beta`` with E, Var over every axis but axis 1.

    Signature and graph behaviour follow ``mygrad.nnet.layers.batchnorm`` (batchnorm.py:114-175);
    ``dp`` (a ``mygrad_b200.dist.DataParallel``) is the one addition: statistics over all ranks.
    """
    if gamma is None:
        gamma = np.array([])
    if beta is None:
        beta = np.array([])
    return Tensor._op(BatchNorm, x, gamma, beta, op_kwargs=dict(eps=eps, dp=dp), constant=constant)
