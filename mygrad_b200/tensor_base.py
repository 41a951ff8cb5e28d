"""Tensor: the graph node, holding a DeviceArray instead of a numpy.ndarray.

Only the part of ``mygrad.Tensor`` that the conv/pool hot path exercises is restated
(src/mygrad/tensor_base.py): construction (:749-859), ``_op`` (:1022-1212), ``backward``
(:1232-1340), ``_backward`` (:1342-1378), ``null_grad``, ``clear_graph`` (:1461-1498) and
``.grad`` (:861-948, non-view branch).  The semantics kept:

* ``_op`` instantiates the Operation, runs it, nulls the grads of its input tensors,
  resolves ``constant`` and records the op (weakly) in each input's ``_ops``;
* ``backward`` topologically sorts the graph (nulling every grad on the way), seeds the
  terminal grad with ones (or the cast/broadcast user grad), runs ``_backward`` in
  order, then clears the graph; intermediate tensors keep their ``.grad``;
* a non-scalar terminal therefore behaves like ``.sum().backward()``.

Operator overloads, in-place ops, views and the NumPy-protocol overrides are out of
scope.
"""
from __future__ import annotations

from collections import deque
from typing import Deque, Optional, Sequence, Set, Tuple, Type
from weakref import ReferenceType

import numpy as np

from . import _graph_tracking as _track
from .device_array import DeviceArray, asdevice, full_like
from .operation_base import Operation

__all__ = ["Tensor", "tensor", "asarray"]


def _broadcast_to(g: DeviceArray, shape: Tuple[int, ...]) -> DeviceArray:
    """materialised np.broadcast_to on the device: zeros(shape) + g"""
    import ctypes as C

    from ._lib import check, lib
    from .device_array import current_stream, zeros

    nd = len(shape)
    i64 = lambda v: (C.c_int64 * max(1, len(v)))(*[int(i) for i in v])
    base, out = zeros(shape, g.dtype), DeviceArray(shape, g.dtype)
    gs = (1,) * (nd - g.ndim) + tuple(g.shape)
    check(lib.mgb_add_broadcast(out.code, nd, i64(shape), i64(shape), i64(gs), C.c_void_p(base.ptr),
                                C.c_void_p(g.ptr), C.c_void_p(out.ptr), C.c_void_p(current_stream())))
    return out


def asarray(x) -> DeviceArray:
    """the DeviceArray behind a Tensor, or `x` moved/viewed onto the device"""
    return x.data if isinstance(x, Tensor) else asdevice(x)


def _collect_and_clear_grads(t: "Tensor", seen: Set[int], order: Deque["Tensor"]) -> None:
    """reverse-topological order of the non-constant graph above `t`; every visited grad
    is nulled (src/mygrad/_utils/__init__.py:36-73).  Iterative so deep graphs are fine."""
    stack = [(t, False)]
    on_path: Set[int] = set()
    while stack:
        node, expanded = stack.pop()
        if expanded:
            on_path.discard(id(node))
            if id(node) not in seen:
                seen.add(id(node))
                order.appendleft(node)
            continue
        node._grad = None
        if node.constant or id(node) in seen:
            continue
        assert id(node) not in on_path, "Computational graph contains a cycle"
        on_path.add(id(node))
        stack.append((node, True))
        if node.creator is not None:
            # reversed: the first variable must be expanded first (pop order)
            for parent in reversed(node.creator.variables):
                stack.append((parent, False))


class Tensor:
    """An N-dimensional device array that records the operations applied to it."""

    __array_priority__ = 15.0

    def __init__(self, x, *, dtype=None, constant: Optional[bool] = None, copy: bool = True,
                 _creator: Optional[Operation] = None):
        if constant is not None and not isinstance(constant, bool):
            raise TypeError(f"`constant` must be a boolean value, got: {constant}")
        self._creator: Optional[Operation] = _creator

        if isinstance(x, Tensor):
            x = x.data
        if isinstance(x, DeviceArray):
            data = x.copy() if copy else x
            if dtype is not None and np.dtype(dtype) != data.dtype:
                data = data.astype(dtype)
        elif hasattr(x, "__cuda_array_interface__"):
            data = asdevice(x, dtype)
            if copy:
                data = data.copy()
        else:
            host = np.asarray(x, dtype=dtype)
            if not np.issubdtype(host.dtype, np.floating):
                raise TypeError(
                    "the device path holds float32/float64 data only; "
                    f"received {host.dtype} (integer tensors are out of scope)"
                )
            data = asdevice(host)  # the upload is the copy
        self.data: DeviceArray = data

        if constant is None:
            constant = False  # float data defaults to non-constant (tensor_base.py:845-846)
        self._constant = constant
        self._grad: Optional[DeviceArray] = None
        self._ops: Set[ReferenceType] = set()

    # ---- metadata ------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return self.data.shape

    @property
    def ndim(self) -> int:
        return self.data.ndim

    @property
    def size(self) -> int:
        return self.data.size

    @property
    def dtype(self) -> np.dtype:
        return self.data.dtype

    @property
    def constant(self) -> bool:
        return self._constant

    @property
    def creator(self) -> Optional[Operation]:
        return self._creator

    @property
    def grad(self) -> Optional[DeviceArray]:
        return self._grad

    def __len__(self) -> int:
        return len(self.data)

    def __repr__(self) -> str:
        return f"Tensor(shape={self.shape}, dtype={self.dtype.name}, constant={self.constant})"

    def numpy(self) -> np.ndarray:
        """host copy of the data"""
        return self.data.get()

    def __array__(self, dtype=None, copy=None):
        return self.data.__array__(dtype)

    def item(self):
        return self.data.item()

    # ---- graph construction --------------------------------------------------------
    @classmethod
    def _op(cls, Op: Type[Operation], *input_vars, op_args: Optional[Sequence] = None,
            op_kwargs: Optional[dict] = None, constant: Optional[bool] = None) -> "Tensor":
        """Runs ``Op`` forward and wires its output into the graph (tensor_base.py:1022-1212)."""
        tensor_vars = tuple(
            var if isinstance(var, Tensor) else cls(var, constant=True, copy=False)
            for var in input_vars
        )
        op_args = tuple(op_args) if op_args is not None else tuple()
        op_kwargs = dict(op_kwargs) if op_kwargs is not None else {}

        f = Op()
        op_out: DeviceArray = f(*tensor_vars, *op_args, **op_kwargs)

        if not _track.TRACK_GRAPH:
            return cls(op_out, constant=constant, copy=False, _creator=None)

        # non-view ops null the grads of their inputs; an op whose output is a view of one of its inputs
        # (Reshape) leaves them alone (tensor_base.py:1134-1172)
        is_view = False
        if f.can_return_view and op_out.base is not None:
            for v in tensor_vars:
                if op_out.base is v.data or op_out.base is v.data.base or op_out is v.data:
                    is_view = True
                    break
        if not is_view:
            for v in input_vars:
                if isinstance(v, Tensor):
                    v._grad = None

        if constant is None:
            constant = None if any(not v.constant for v in tensor_vars) else True

        ref_f = ReferenceType(f)
        for v in tensor_vars:
            v._ops.add(ref_f)

        return cls(op_out, constant=constant, copy=False, _creator=f)

    # ---- back-propagation ----------------------------------------------------------
    def backward(self, grad=None) -> None:
        """Computes d(self)/d(t) for every non-constant tensor t above `self` (tensor_base.py:1232-1340)."""
        if not _track.TRACK_GRAPH:
            return
        if self.constant:
            self.clear_graph()
            return

        order: Deque[Tensor] = deque()
        _collect_and_clear_grads(self, set(), order)

        if grad is not None:
            g = grad.data if isinstance(grad, Tensor) else grad
            if not isinstance(g, DeviceArray):
                if hasattr(g, "__cuda_array_interface__"):
                    g = asdevice(g, self.dtype)
                else:
                    host = np.asarray(g, dtype=self.dtype)
                    if host.shape != self.shape:
                        try:
                            host = np.broadcast_to(host, self.shape)
                        except ValueError:
                            raise ValueError(
                                f"`tensor.backward(grad)` was passed a gradient with an incompatible "
                                f"shape.\n`grad` must be broadcast-compatible with "
                                f"`tensor.shape={self.shape}`\nGot `grad.shape={host.shape}`"
                            )
                    # (np.ascontiguousarray would turn a 0-d gradient into a 1-d one)
                    g = asdevice(np.array(host, dtype=self.dtype, order="C", copy=True))
            if g.dtype != self.dtype:
                g = g.astype(self.dtype)
            if g.shape != self.shape:
                # a device gradient of broadcast-compatible shape is broadcast like a host one
                # (tensor_base.py:1311-1330 uses np.broadcast_to): zeros(shape) + g on the device
                try:
                    ok = np.broadcast_shapes(g.shape, self.shape) == self.shape
                except ValueError:
                    ok = False
                if not ok:
                    raise ValueError(
                        f"`tensor.backward(grad)` was passed a gradient with an incompatible shape.\n"
                        f"`grad` must be broadcast-compatible with `tensor.shape={self.shape}`\n"
                        f"Got `grad.shape={g.shape}`"
                    )
                g = _broadcast_to(g, self.shape)
            seed = g
        else:
            seed = full_like(self.data, 1.0)
        self._grad = seed

        if self.creator is not None:
            for t in order:
                t._backward()
        self.clear_graph()

    def _backward(self) -> None:
        assert self._grad is not None, (
            f"backprop, post grad-accumulation, was triggered on a tensor with no gradient\n{self}"
        )
        assert self._grad.shape == self.shape, (
            f"A tensor and its associated gradient must possess the same shape. Got:"
            f"\ntensor-shape: {self.shape}\ngrad-shape: {self._grad.shape}"
        )
        if self._creator is not None:
            self._creator.backward(self._grad)

    # ---- the little operator sugar config 5 uses (everything else is out of scope) -----
    def __add__(self, other) -> "Tensor":
        from .math_ops import add

        return add(self, other)

    def __radd__(self, other) -> "Tensor":
        from .math_ops import add

        return add(other, self)

    def __matmul__(self, other) -> "Tensor":
        from .math_ops import matmul

        return matmul(self, other)

    def reshape(self, *newshape) -> "Tensor":
        from .math_ops import reshape

        return reshape(self, *newshape)

    def null_grad(self) -> "Tensor":
        self._grad = None
        return self

    def clear_graph(self) -> None:
        """Drops creators and op references of this tensor and everything above it."""
        stack = [self]
        while stack:
            t = stack.pop()
            t._ops.clear()
            creator = t._creator
            if creator is None:
                continue
            t._creator = None
            stack.extend(creator.variables)


def tensor(x, *, dtype=None, constant: Optional[bool] = None, copy: bool = True) -> Tensor:
    return Tensor(x, dtype=dtype, constant=constant, copy=copy)
