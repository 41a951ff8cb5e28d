"""conv_nd on B200: N-d cross-correlation forward, dX and dW as implicit-GEMM kernels.

Drop-in for ``mygrad.nnet.layers.conv_nd`` / ``ConvND``
(src/mygrad/nnet/layers/conv.py:14-155 op, :158-395 wrapper): same signature, same
attributes set on the op (``variables, padding, stride, dilation``), same exception
types (AssertionError for malformed stride/padding/dilation, ValueError for
incompatible shapes), same output layout ``(N, F, G0, ...)``.
"""
from __future__ import annotations

import ctypes as C
from numbers import Integral
from typing import Optional, Tuple, Union

import numpy as np

from ..._lib import ConvDesc, check, lib
from ...device_array import DeviceArray, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor

__all__ = ["conv_nd", "ConvND"]

_MAX_SPATIAL = 3

# One grow-only scratch buffer per stream, shared by every conv call on that stream: the
# kernels of consecutive calls are ordered on the stream, so the buffer can be re-used as
# soon as the next call is enqueued.  (Allocating multi-GB scratch per call makes the
# stream-ordered pool grow and trim, which costs tens of milliseconds when it happens.)
_WORKSPACES: dict = {}


def _ptr(a: Optional[DeviceArray]) -> C.c_void_p:
    return C.c_void_p(a.ptr if a is not None else 0)


def _scratch(nbytes: int) -> DeviceArray:
    key = current_stream()
    buf = _WORKSPACES.get(key)
    if buf is None or buf.nbytes < nbytes:
        _WORKSPACES[key] = None  # drop the old buffer before asking for a bigger one
        buf = DeviceArray(((nbytes * 5 // 4 + 7) // 8,), np.float64)
        _WORKSPACES[key] = buf
    return buf


def _as_tuple(value, n: int, *, minimum: int, name: str) -> Tuple[int, ...]:
    """int -> n-tuple; sequences must have length n and integer entries >= minimum.
    Violations are AssertionErrors, as in the reference (conv.py:35-61)."""
    if isinstance(value, Integral):
        out = (int(value),) * n
        ok = value >= minimum
    else:
        seq = list(value)
        ok = len(seq) == n and all(isinstance(v, Integral) and v >= minimum for v in seq)
        out = tuple(int(v) for v in seq) if ok else ()
    assert ok, f"`{name}` must be an integer or a length-{n} sequence of integers >= {minimum}"
    return out


class ConvND(Operation):
    def __call__(self, x: Tensor, w: Tensor, *, stride, padding=0, dilation=1) -> DeviceArray:
        out_shape = self._plan(x, w, stride=stride, padding=padding, dilation=dilation)
        xin, win = self._operands()
        y = DeviceArray(out_shape, self._dtype)
        ws, ws_bytes = self._workspace(0)
        # operand cache of the tensor-core path: x planes are re-used by dW, dy planes by dX and dW
        self._x_planes = self._make_planes(0, xin)
        self._dy_planes = None
        check(lib.mgb_conv_fwd_p(C.byref(self._desc), C.c_void_p(xin.ptr), C.c_void_p(win.ptr),
                                 C.c_void_p(y.ptr), _ptr(self._x_planes), C.c_void_p(ws.ptr if ws else 0),
                                 C.c_size_t(ws_bytes), C.c_void_p(current_stream())))
        return y

    def _plan(self, x: Tensor, w: Tensor, *, stride, padding=0, dilation=1) -> Tuple[int, ...]:
        """validation, the op's public attributes and the C descriptor; returns the output shape"""
        self.variables = (x, w)
        xd, wd = x.data, w.data  # (N, C, X0, ...), (F, C, W0, ...)

        assert xd.ndim > 2
        assert xd.ndim == wd.ndim
        assert wd.shape[1] == xd.shape[1], "The channel-depth of the batch and filters must agree"

        nd = wd.ndim - 2
        x_shape, w_shape = xd.shape[2:], wd.shape[2:]
        dilation = _as_tuple(dilation, nd, minimum=1, name="dilation")
        padding = _as_tuple(padding, nd, minimum=0, name="padding")
        stride = _as_tuple(stride, nd, minimum=1, name="stride")

        out_shape = []
        ok = True
        for X, W, s, p, d in zip(x_shape, w_shape, stride, padding, dilation):
            num = X + 2 * p - ((W - 1) * d + 1)
            ok = ok and num >= 0 and num % s == 0
            out_shape.append(num // s + 1)
        if not ok:
            raise ValueError(
                "Stride and kernel dimensions are incompatible: \n"
                f"Input dimensions: {tuple(x_shape)}\n"
                f"Stride dimensions: {tuple(stride)}\n"
                f"Kernel dimensions: {tuple(w_shape)}\n"
                f"Padding dimensions: {tuple(padding)}\n"
                f"Dilation dimensions: {tuple(dilation)}\n"
            )
        if nd > _MAX_SPATIAL:
            raise NotImplementedError(
                f"the B200 path convolves over at most {_MAX_SPATIAL} spatial dims, got {nd}"
            )

        self.padding = np.array(padding, dtype=int)
        self.stride = np.array(stride, dtype=int)
        self.dilation = np.array(dilation, dtype=int)

        # NumPy promotion of (x, w): f32*f32 -> f32, anything with f64 -> f64
        self._dtype = np.result_type(xd.dtype, wd.dtype)
        self._desc = ConvDesc()
        d = self._desc
        d.dtype = 0 if self._dtype == np.float32 else 1
        d.nd = nd
        d.N, d.C, d.F = xd.shape[0], xd.shape[1], wd.shape[0]
        for i in range(nd):
            d.X[i], d.W[i] = x_shape[i], w_shape[i]
            d.stride[i], d.pad[i], d.dil[i] = stride[i], padding[i], dilation[i]
        self._x_planes = None
        self._dy_planes = None
        return (xd.shape[0], wd.shape[0], *out_shape)

    def _make_planes(self, tensor: int, src: DeviceArray) -> Optional[DeviceArray]:
        n = C.c_size_t(0)
        check(lib.mgb_conv_planes_bytes(C.byref(self._desc), tensor, C.byref(n)))
        if n.value == 0:
            return None
        planes = DeviceArray(((n.value + 7) // 8,), np.float64)
        check(lib.mgb_conv_make_planes(C.byref(self._desc), tensor, C.c_void_p(src.ptr),
                                       C.c_void_p(planes.ptr), C.c_void_p(current_stream())))
        return planes

    # operands in the compute dtype (mixed f32/f64 inputs are promoted, as NumPy does)
    def _operands(self) -> Tuple[DeviceArray, DeviceArray]:
        x, w = (v.data for v in self.variables)
        return x.astype(self._dtype, copy=False), w.astype(self._dtype, copy=False)

    def _workspace(self, which: int):
        n = C.c_size_t(0)
        check(lib.mgb_conv_workspace_bytes(C.byref(self._desc), which, C.byref(n)))
        if n.value == 0:
            return None, 0
        return _scratch(n.value), n.value

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        g = grad.astype(self._dtype, copy=False)
        if index == 0:  # dX, shape of x, dtype of x (conv.py:118)
            # planes of exactly this gradient may already exist: MaxPoolND.backward_var writes them together
            # with dy when it feeds a convolution (pooling.py: the transparent half of the conv -> pool fusion)
            cached = self._dy_planes
            if cached is None or cached[0] != g.ptr:
                self._dy_planes = (g.ptr, self._make_planes(1, g))
            return self._dgrad(g, self._dy_planes[1])
        # dW: dy planes made by the dX pass of the same backward call are valid for this grad only
        cached = self._dy_planes
        dyp = cached[1] if cached is not None and cached[0] == g.ptr else None
        self._dy_planes = None
        return self._wgrad(g, dyp)

    def _dgrad(self, g: Optional[DeviceArray], dy_planes: Optional[DeviceArray]) -> DeviceArray:
        """dX from dy (`g`) and/or its channels-last planes (the fused conv_pool op passes planes only)"""
        x = self.variables[0].data
        _, win = self._operands()
        dx = DeviceArray(x.shape, self._dtype)
        ws, ws_bytes = self._workspace(1)
        check(lib.mgb_conv_dgrad_p(C.byref(self._desc), _ptr(g), C.c_void_p(win.ptr),
                                   C.c_void_p(dx.ptr), _ptr(dy_planes),
                                   C.c_void_p(ws.ptr if ws else 0), C.c_size_t(ws_bytes),
                                   C.c_void_p(current_stream())))
        return dx.astype(x.dtype, copy=False)

    def _wgrad(self, g: Optional[DeviceArray], dy_planes: Optional[DeviceArray]) -> DeviceArray:
        w = self.variables[1].data
        xin, _ = self._operands()
        dw = DeviceArray(w.shape, self._dtype)
        ws, ws_bytes = self._workspace(2)
        # inside DataParallel.fused_wgrad() the split-K reduction also sums over the ranks
        # (one kernel over NVLink peer memory); otherwise dW is this rank's own
        from ...dist import fused_dp

        dp = fused_dp()
        fused = dp is not None and dp.fuses(dw.nbytes)
        wgrad = lib.mgb_conv_wgrad_dp if fused else lib.mgb_conv_wgrad_p
        dw.reduced = fused          # already the sum over the ranks: allreduce() will only scale it
        check(wgrad(C.byref(self._desc), _ptr(g), C.c_void_p(xin.ptr),
                    C.c_void_p(dw.ptr), _ptr(dy_planes), _ptr(self._x_planes),
                    C.c_void_p(ws.ptr if ws else 0), C.c_size_t(ws_bytes), C.c_void_p(current_stream())))
        return dw


def conv_nd(
    x,
    filter_bank,
    *,
    stride: Union[int, Tuple[int, ...]],
    padding: Union[int, Tuple[int, ...]] = 0,
    dilation: Union[int, Tuple[int, ...]] = 1,
    constant: Optional[bool] = None,
) -> Tensor:
    """N-dimensional cross-correlation of the batch `x` (N, C, X0, ...) with the filter
    bank (F, C, W0, ...) -> (N, F, G0, ...), G = (X + 2p - ((W-1)d + 1))/s + 1.

    Signature, validation and graph behaviour follow ``mygrad.nnet.layers.conv_nd``
    (conv.py:158-166, 373-395); the arithmetic runs on the GPU.
    """
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
        ConvND,
        x,
        filter_bank,
        op_kwargs={"stride": stride, "padding": padding, "dilation": dilation},
        constant=constant,
    )
