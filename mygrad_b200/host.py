"""Host-buffer entry point: conv_nd -> max_pool forward + backward with NumPy in, NumPy out.

A user of the reference holds ``numpy`` arrays on the host.  ``HostLayerPipeline.step`` is
the call that keeps that contract — ``x, w`` in host memory, ``out, x.grad, w.grad`` back in
host memory — while hiding most of the PCIe time: the batch axis is cut into chunks (every
output sample depends only on its own input sample; dW is additive over the
batch) and three streams run H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1
concurrently.  The arithmetic is the same ``conv_nd`` / ``max_pool`` / ``Tensor.backward``
path as everywhere else; only copies and stream ordering live here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from ._lib import check, lib
from .device_array import DeviceArray, current_stream, empty, pinned_empty
from .dist import shard_bounds
from .nnet.layers import conv_nd, max_pool
from .tensor_base import Tensor

__all__ = ["HostLayerPipeline"]


def _stream() -> C.c_void_p:
    s = C.c_void_p()
    check(lib.mgb_stream_create(C.byref(s)))
    return s


def _event() -> C.c_void_p:
    e = C.c_void_p()
    check(lib.mgb_event_create(C.byref(e)))
    return e


class HostLayerPipeline:
    """``out = max_pool(conv_nd(x, w, ...), pool, pool_stride); out.backward()`` on host arrays.

    Page-locked host arrays (``mygrad_b200.pinned_empty``) make the copies asynchronous;
    pageable arrays work but serialise.  Results: ``(out, dx, dw)`` NumPy arrays (views of the
    pipeline's pinned output buffers unless ``out= / dx= / dw=`` are given).
    """

    def __init__(self, *, stride, padding=0, dilation=1, pool: Optional[Sequence[int]] = (2, 2),
                 pool_stride=2, chunks: int = 8):
        self.conv_kw = dict(stride=stride, padding=padding, dilation=dilation)
        self.pool, self.pool_stride, self.chunks = pool, pool_stride, max(1, int(chunks))
        self._h2d, self._d2h = _stream(), _stream()
        self._ev: dict = {}
        self._bufs: dict = {}

    def _event_for(self, kind: str, i: int) -> C.c_void_p:
        key = (kind, i)
        if key not in self._ev:
            self._ev[key] = _event()
        return self._ev[key]

    def _buffer(self, name: str, shape, dtype, pinned=False):
        key = (name, tuple(shape), np.dtype(dtype).str)
        if key not in self._bufs:
            self._bufs[key] = pinned_empty(shape, dtype) if pinned else empty(shape, dtype)
        return self._bufs[key]

    def step(self, x: np.ndarray, w: np.ndarray, *, out: Optional[np.ndarray] = None,
             dx: Optional[np.ndarray] = None, dw: Optional[np.ndarray] = None
             ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        compute = C.c_void_p(current_stream())
        x, w = self._check_inputs(x, w)
        n = x.shape[0]
        nchunks = min(self.chunks, max(1, n))
        bounds = [shard_bounds(n, i, nchunks) for i in range(nchunks)]
        dt = x.dtype
        # NumPy promotion of (x, w) is the dtype of y / out / dW (conv.py:97-99); x.grad keeps x's dtype (:118)
        odt = np.result_type(x.dtype, w.dtype)
        out_shape = self._out_shape(x.shape, w.shape)
        out = self._check_result("out", out, out_shape, odt)
        dx = self._check_result("dx", dx, x.shape, dt)
        dw = self._check_result("dw", dw, w.shape, odt)
        # the filter bank goes first, on the compute stream
        wd = self._buffer("w", w.shape, w.dtype)
        check(lib.mgb_memcpy_h2d(C.c_void_p(wd.ptr), C.c_void_p(w.ctypes.data), C.c_size_t(w.nbytes), compute))
        dw_total = None
        keep = []            # device results stay referenced until every copy has finished
        per = max(hi - lo for lo, hi in bounds)
        xin = [self._buffer(f"x{j}", (per, *x.shape[1:]), dt) for j in range(2)]
        sample_bytes = x.nbytes // max(1, n)
        for i, (lo, hi) in enumerate(bounds):
            m = hi - lo
            xd_full = xin[i % 2]
            # the input slot was last read by the kernels of chunk i-2
            if i >= 2:
                check(lib.mgb_stream_wait_event(self._h2d, self._event_for("comp", i - 2)))
            check(lib.mgb_memcpy_h2d(C.c_void_p(xd_full.ptr), C.c_void_p(x.ctypes.data + lo * sample_bytes),
                                     C.c_size_t(m * sample_bytes), self._h2d))
            check(lib.mgb_event_record(self._event_for("h2d", i), self._h2d))
            check(lib.mgb_stream_wait_event(compute, self._event_for("h2d", i)))
            xd = xd_full if m == per else DeviceArray((m, *x.shape[1:]), dt, _ptr=xd_full.ptr, _base=xd_full)
            xt, wt = Tensor(xd, copy=False), Tensor(wd, copy=False)
            y = conv_nd(xt, wt, **self.conv_kw)
            o = max_pool(y, self.pool, self.pool_stride) if self.pool else y
            o.backward()
            if dw_total is None:
                dw_total = wt.grad
            else:
                dw_total += wt.grad
            check(lib.mgb_event_record(self._event_for("comp", i), compute))
            if o.shape[1:] != out_shape[1:] or o.data.dtype != odt or xt.grad.dtype != dt:
                raise RuntimeError(f"internal: result {o.shape} {o.data.dtype} does not match the planned "
                                   f"{out_shape} {odt}")
            check(lib.mgb_stream_wait_event(self._d2h, self._event_for("comp", i)))
            # o.data / xt.grad are read by the copy stream: `keep` holds them until BOTH streams have been
            # synchronised below, so their blocks return to the allocation cache only after the copies are done
            # (buffers handed to a side stream without such a fence need DeviceArray.record_stream)
            o_bytes = o.data.nbytes // max(1, m)
            check(lib.mgb_memcpy_d2h(C.c_void_p(out.ctypes.data + lo * o_bytes), C.c_void_p(o.data.ptr),
                                     C.c_size_t(m * o_bytes), self._d2h))
            check(lib.mgb_memcpy_d2h(C.c_void_p(dx.ctypes.data + lo * sample_bytes), C.c_void_p(xt.grad.ptr),
                                     C.c_size_t(m * sample_bytes), self._d2h))
            keep.append((o, xt, wt, y))
        # dW is complete once the last chunk's accumulation has run: copy it on the compute stream
        check(lib.mgb_memcpy_d2h(C.c_void_p(dw.ctypes.data), C.c_void_p(dw_total.ptr), C.c_size_t(dw.nbytes),
                                 compute))
        check(lib.mgb_stream_synchronize(compute))
        check(lib.mgb_stream_synchronize(self._d2h))
        del keep
        return out, dx, dw

    # ---- argument checking: raw pointers go to the copy engines, so nothing unchecked may reach them ----
    @staticmethod
    def _check_inputs(x: np.ndarray, w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        x, w = np.asarray(x), np.asarray(w)
        for name, a in (("x", x), ("w", w)):
            if a.dtype not in (np.float32, np.float64):
                raise TypeError(f"`{name}` must be float32 or float64, got {a.dtype}")
        if x.ndim < 3 or x.ndim != w.ndim or x.shape[1] != w.shape[1]:
            raise ValueError(f"`x` {x.shape} and `w` {w.shape} are not a (N, C, X...) batch and a (F, C, W...) filter bank")
        # sliced / transposed views would upload the wrong bytes: make them contiguous (a copy, like the
        # reference's own np.ascontiguousarray in sliding_window_view, nnet/layers/utils.py:198-199)
        return np.ascontiguousarray(x), np.ascontiguousarray(w)

    def _out_shape(self, x_shape, w_shape):
        nd = len(x_shape) - 2
        tup = lambda v: (int(v),) * nd if np.isscalar(v) else tuple(int(i) for i in v)
        s, p, d = tup(self.conv_kw["stride"]), tup(self.conv_kw["padding"]), tup(self.conv_kw["dilation"])
        g = []
        for X, W, si, pi, di in zip(x_shape[2:], w_shape[2:], s, p, d):
            num = X + 2 * pi - ((W - 1) * di + 1)
            if num < 0 or num % si:
                raise ValueError("Stride and kernel dimensions are incompatible")
            g.append(num // si + 1)
        if self.pool:
            ps = tup(self.pool_stride) if not np.isscalar(self.pool_stride) else (int(self.pool_stride),) * len(self.pool)
            k = len(g) - len(self.pool)
            for i, (P, st) in enumerate(zip(self.pool, ps)):
                num = g[k + i] - P
                if num < 0 or num % st:
                    raise ValueError("Stride and kernel dimensions are incompatible")
                g[k + i] = num // st + 1
        return (x_shape[0], w_shape[0], *g)

    def _check_result(self, name: str, arr: Optional[np.ndarray], shape, dtype) -> np.ndarray:
        if arr is None:
            return self._buffer(name, shape, dtype, pinned=True)
        if (not isinstance(arr, np.ndarray) or tuple(arr.shape) != tuple(shape) or arr.dtype != np.dtype(dtype)
                or not arr.flags["C_CONTIGUOUS"] or not arr.flags["WRITEABLE"]):
            raise ValueError(f"`{name}=` must be a writeable C-contiguous {np.dtype(dtype).name} array of shape "
                             f"{tuple(shape)}; got {getattr(arr, 'dtype', type(arr))} {getattr(arr, 'shape', '')}")
        return arr

    def close(self) -> None:
        for e in self._ev.values():
            lib.mgb_event_destroy(e)
        self._ev.clear()
        for s in (self._h2d, self._d2h):
            if s is not None:
                lib.mgb_stream_destroy(s)
        self._h2d = self._d2h = None
        self._bufs.clear()
