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
        # checks in the order of check_loss_inputs (nnet/losses/_utils.py:6-48): integer dtype (TypeError), then
        # shape (ValueError).  Device-resident labels can only be float-typed here (an extension); they must
        # hold whole numbers.
        if isinstance(y_true, DeviceArray):
            labels_host = y_true.get()
            if not np.all(labels_host == np.round(labels_host)):
                raise TypeError("`y_true` must hold integer class indices")
            labels_host = labels_host.astype(np.int64)
        else:
            labels_host = np.asarray(y_true)
            if not np.issubdtype(labels_host.dtype, np.integer):
                raise TypeError(f"`y_true` must be an integer-type array-like object, got {labels_host.dtype}")
        if labels_host.ndim != 1 or labels_host.shape[0] != scores.shape[0]:
            raise ValueError(
                "`y_true` must be a shape-(N,) array: \n"
                f"\tExpected shape-{(scores.shape[0],)}\n\tGot shape-{labels_host.shape}"
            )
        # the reference gathers x[range(N), y_true] (softmax_crossentropy.py:41): negative indices count from
        # the end, anything outside [-C, C) is an IndexError
        ncls = scores.shape[1]
        if labels_host.size and (labels_host.min() < -ncls or labels_host.max() >= ncls):
            bad = labels_host[(labels_host < -ncls) | (labels_host >= ncls)][0]
            raise IndexError(f"index {bad} is out of bounds for axis 1 with size {ncls}")
        labels_host = np.where(labels_host < 0, labels_host + ncls, labels_host)
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
