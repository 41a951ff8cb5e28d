"""sliding_window_view for device arrays.

``mygrad.nnet.layers.utils.sliding_window_view(arr, window_shape, step, dilation=None)``
(src/mygrad/nnet/layers/utils.py:7-228) returns a read-only STRIDED VIEW of shape
``(G..., lead..., W...)``.  A ``DeviceArray`` is C-contiguous by construction, so the device
version returns the same values as a freshly gathered array (one bandwidth-bound kernel).
The convolution and pooling kernels never build this array — they index the windows in place;
this function exists for users of the public utility.  Argument validation and the exception
types follow the reference line by line (utils.py:123-193).
"""
from __future__ import annotations

import ctypes as C
from numbers import Integral

import numpy as np

from ..._lib import check, lib
from ...device_array import DeviceArray, asdevice, current_stream

__all__ = ["sliding_window_view"]


def sliding_window_view(arr, window_shape, step, dilation=None) -> DeviceArray:
    arr = asdevice(arr)
    if not hasattr(window_shape, "__iter__"):
        raise TypeError(f"`window_shape` must be a sequence of positive integers, got: {window_shape}")
    window_shape = tuple(window_shape)
    if not all(isinstance(i, Integral) and i > 0 for i in window_shape):
        raise TypeError(f"`window_shape` must be a sequence of positive integers, got: {window_shape}")
    if len(window_shape) > arr.ndim:
        raise ValueError(
            f"`window_shape` ({window_shape}) cannot specify more values than `arr.ndim` ({arr.ndim})."
        )
    if not isinstance(step, Integral) and not hasattr(step, "__iter__"):
        raise TypeError(f"`step` must be a positive integer or a sequence of positive integers, got: {step}")
    step = (int(step),) * len(window_shape) if isinstance(step, Integral) else tuple(step)
    if not all(isinstance(i, Integral) and i > 0 for i in step):
        raise ValueError(f"`step` must be a positive integer or a sequence of positive integers, got: {step}")
    if any(i > j for i, j in zip(window_shape[::-1], arr.shape[::-1])):
        raise ValueError(
            "Each size of the window-shape must fit within the trailing dimensions of `arr`."
            f"{window_shape} does not fit in {arr.shape[-len(window_shape):]}"
        )
    if dilation is not None and not isinstance(dilation, Integral) and not hasattr(dilation, "__iter__"):
        raise TypeError(
            f"`dilation` must be None, a positive integer, or a sequence of positive integers, got: {dilation}"
        )
    if dilation is None:
        dilation = (1,) * len(window_shape)
    else:
        dilation = (int(dilation),) * len(window_shape) if isinstance(dilation, Integral) else tuple(dilation)
        if not all(isinstance(i, Integral) and i > 0 for i in dilation) or len(dilation) != len(window_shape):
            raise ValueError(
                "`dilation` must be None, a positive integer, or a sequence of positive integers with the "
                f"same length as `window_shape` ({window_shape}), got: {dilation}"
            )
        if any(w * d > s for w, d, s in zip(window_shape[::-1], dilation[::-1], arr.shape[::-1])):
            raise ValueError(
                f"The dilated window ({tuple(w * d for w, d in zip(window_shape, dilation))}) must fit within "
                f"the trailing dimensions of `arr` ({arr.shape[-len(window_shape):]})"
            )
    nd = len(window_shape)
    if nd < 1 or nd > 3:
        raise NotImplementedError(f"the device path windows 1 to 3 trailing axes, got {nd}")
    if len(step) != nd:
        raise ValueError(f"`step` must have one entry per window axis ({nd}), got: {step}")
    lead_shape, x_shape = arr.shape[: arr.ndim - nd], arr.shape[arr.ndim - nd:]
    out_sp = tuple((x - ((w - 1) * d + 1)) // s + 1 for x, w, s, d in zip(x_shape, window_shape, step, dilation))
    out = DeviceArray((*out_sp, *lead_shape, *window_shape), arr.dtype)
    i64 = lambda v: (C.c_int64 * nd)(*[int(i) for i in v])
    check(lib.mgb_window_gather(arr.code, nd, int(np.prod(lead_shape, dtype=np.int64)), i64(x_shape),
                                i64(window_shape), i64(step), i64(dilation), C.c_void_p(arr.ptr),
                                C.c_void_p(out.ptr), C.c_void_p(current_stream())))
    return out
