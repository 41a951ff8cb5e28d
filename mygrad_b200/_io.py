"""save / load of device tensors in the reference's on-disk format.

``mygrad.save`` writes ``numpy.savez(file, data=tensor.data[, grad=tensor.grad])`` and
``mygrad.load`` rebuilds the tensor and re-attaches the gradient with ``backward(grad)``
(src/mygrad/_io.py:54-63, 112-125).  The same two keys are used here, so files are
interchangeable with the reference in both directions; the only work specific to this path is
the device<->host copy on either side of NumPy's .npz codec.
"""
from __future__ import annotations

from pathlib import Path
from typing import BinaryIO, Union

import numpy as np

from .tensor_base import Tensor

__all__ = ["save", "load"]

_FileLike = Union[str, Path, BinaryIO]


def save(file: _FileLike, tensor: Tensor) -> None:
    """Saves ``tensor.data`` (key ``data``) and, if present, ``tensor.grad`` (key ``grad``) as .npz."""
    if not isinstance(tensor, Tensor):
        raise TypeError(f"mygrad.save requires a Tensor-type object, got type {type(tensor)}")
    if tensor.grad is not None:
        np.savez(file, data=tensor.data.get(), grad=tensor.grad.get())
    else:
        np.savez(file, data=tensor.data.get())


def load(file: _FileLike) -> Tensor:
    """Loads a tensor written by ``save`` (or by ``mygrad.save``); a saved gradient becomes ``.grad``."""
    loaded = np.load(file)
    tensor = Tensor(loaded["data"])
    if "grad" in loaded:
        tensor.backward(loaded["grad"])
    return tensor
