"""The `Operation` protocol on device buffers — the drop-in boundary.

Mirrors ``mygrad.operation_base.Operation`` (src/mygrad/operation_base.py:54-228):
subclasses implement ``__call__(*tensors, **kw) -> array`` (and must set
``self.variables``) and ``backward_var(grad, index) -> array``; ``backward`` walks the
inputs, un-broadcasts, casts to the variable's dtype and stores or accumulates the
gradient.  Arrays are ``DeviceArray`` instead of ``numpy.ndarray``; the un-broadcast
reduction, the defensive copy, the cast and the ``+=`` are device kernels
(mgb_reduce_broadcast / mgb_copy_cast / mgb_axpy).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from numbers import Real
from typing import TYPE_CHECKING, Any, Dict, Optional, Tuple

import numpy as np

from .device_array import DeviceArray, asdevice
from .errors import InvalidBackprop, InvalidGradient, SkipGradient

if TYPE_CHECKING:  # pragma: no cover
    from .tensor_base import Tensor

__all__ = ["Operation"]


class Operation(ABC):
    """Base class of differentiable operations (see module docstring)."""

    # ConvND / MaxPoolND never return views (operation_base.py:75)
    can_return_view: bool = False

    variables: Tuple["Tensor", ...]

    def __init__(self):
        self.replay_args: Optional[Tuple[Any, ...]] = None
        self.replay_kwargs: Optional[Dict[str, Any]] = None
        self.replay_force_constant: Optional[bool] = None
        self.where = True

    @staticmethod
    def grad_post_process_fn(grad: DeviceArray, var_shape: Tuple[int, ...]) -> DeviceArray:
        """un-broadcast `grad` to `var_shape` (operation_base.py:89-105)"""
        if grad.shape == tuple(var_shape):
            return grad
        if grad.ndim < len(var_shape):
            raise ValueError(
                f"The dimensionality of the gradient of the broadcasted operation ({grad.ndim}) "
                f"is less than that of its associated variable ({len(var_shape)})"
            )
        return grad.sum_to_shape(var_shape)

    @abstractmethod
    def __call__(self, *input_vars: "Tensor", **kwargs) -> DeviceArray:  # pragma: no cover
        raise NotImplementedError()

    @abstractmethod
    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:  # pragma: no cover
        raise NotImplementedError()

    def backward(self, grad: DeviceArray, **kwargs) -> None:
        """Back-propagates `grad` to every non-constant input (operation_base.py:162-228)."""
        for index, var in enumerate(self.variables):
            if var.constant:
                continue

            if not var._ops:
                raise InvalidBackprop(
                    f"Part of the computational graph containing this tensor, {var}, was "
                    f"'cleared' prior to backprop.\nIt is recommended that you clear all "
                    f"computational graphs and restart your computation."
                )

            try:
                backed = self.backward_var(grad, index, **kwargs)
            except SkipGradient:
                continue

            if not isinstance(backed, (DeviceArray, np.ndarray, np.number, Real)):
                raise InvalidGradient(
                    f"An invalid gradient-value was passed to:"
                    f"\n\t`{type(self).__name__}.backward_var(<gradient>, index={index})`"
                    f"\nGradients are expected to be real-valued scalars or "
                    f"arrays, got a gradient of type: {type(backed)}"
                )
            if not isinstance(backed, DeviceArray):
                backed = asdevice(np.asarray(backed))

            backed = self.grad_post_process_fn(backed, var.shape)
            assert backed.shape == var.shape, (backed.shape, var.shape)

            if var._grad is None:
                # a view of `grad`, or `grad` itself, must not become var's grad:
                # later arrivals are accumulated into it in place
                if backed.base is not None or backed is grad:
                    backed = backed.copy()
                if backed.dtype != var.dtype:
                    backed = backed.astype(var.dtype, copy=False)
                var._grad = backed
            else:
                var._grad += backed
