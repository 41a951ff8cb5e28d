"""softmax_crossentropy on the device — drop-in for
``mygrad.nnet.losses.softmax_crossentropy`` (src/mygrad/nnet/losses/softmax_crossentropy.py:14-50):
mean over the batch of ``-log softmax(x)[n, y_n]`` computed through log-sum-exp
(src/mygrad/math/_special.py:4-68); the op caches ``(softmax - onehot)/N`` for the backward."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ... import _graph_tracking as _track
from ..._lib import check, lib
from ...device_array import DeviceArray, asdevice, current_stream
from ...operation_base import Operation
from ...tensor_base import Tensor

__all__ = ["softmax_crossentropy", "SoftmaxCrossEntropy"]


class SoftmaxCrossEntropy(Operation):
    def __call__(self, x: Tensor, y_true) -> DeviceArray:
        if isinstance(y_true, Tensor):
            y_true = y_true.data
        scores = x.data
        # same input checks as nnet/losses/_utils.py::check_loss_inputs
        if scores.ndim != 2:
            raise ValueError(f"`x` must be a 2-dimensional array, got shape {scores.shape}")
        if isinstance(y_true, DeviceArray):
            labels_host = y_true.get()
        else:
            labels_host = np.asarray(y_true)
        if labels_host.ndim != 1 or labels_host.shape[0] != scores.shape[0]:
            raise ValueError(
                f"`y_true` must be a 1-dimensional array of length {scores.shape[0]}, got shape {labels_host.shape}"
            )
        if not (np.issubdtype(labels_host.dtype, np.integer) or np.all(labels_host == np.round(labels_host))):
            raise TypeError("`y_true` must hold integer class indices")
        if labels_host.size and (labels_host.min() < 0 or labels_host.max() >= scores.shape[1]):
            raise ValueError("class indices in `y_true` must lie in [0, C)")
        self.variables = (x,)
        labels = asdevice(np.ascontiguousarray(labels_host, dtype=scores.dtype))
        loss = DeviceArray((), scores.dtype)
        self.back = DeviceArray(scores.shape, scores.dtype) if _track.TRACK_GRAPH else None
        check(lib.mgb_softmax_xent(scores.code, scores.shape[0], scores.shape[1], C.c_void_p(scores.ptr),
                                   C.c_void_p(labels.ptr), C.c_void_p(loss.ptr),
                                   C.c_void_p(self.back.ptr if self.back is not None else 0),
                                   C.c_void_p(current_stream())))
        return loss

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        g = grad.astype(self.back.dtype, copy=False)
        out = DeviceArray(self.back.shape, self.back.dtype)
        check(lib.mgb_scale_by(out.code, C.c_size_t(out.size), C.c_void_p(g.ptr), C.c_void_p(self.back.ptr),
                               C.c_void_p(out.ptr), C.c_void_p(current_stream())))
        return out


def softmax_crossentropy(x, y_true, *, constant: Optional[bool] = None) -> Tensor:
    """average softmax cross-entropy of the (N, C) scores `x` against the N class indices `y_true`"""
    return Tensor._op(SoftmaxCrossEntropy, x, op_args=(y_true,), constant=constant)
