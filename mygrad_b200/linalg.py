"""``einsum`` on device tensors.

Mirrors ``mygrad.einsum`` / ``EinSum`` (src/mygrad/linalg/funcs.py:178-440, linalg/ops.py:79-243):
NumPy's subscript language — explicit and implicit output, repeated labels inside an operand (traces
and diagonals), ellipsis broadcasting, the sublist call form — with back-propagation through every
operand, one gradient per distinct (tensor, subscript) pair scaled by its multiplicity
(ops.py:118-150), gradients of traced operands written along the diagonals of a zero array
(ops.py:215-236) and labels that were summed without a partner broadcast back (ops.py:168-190).

The reference delegates the arithmetic to ``numpy.einsum``; here every contraction, forward and
backward, is ONE launch of ``mgb_einsum`` (csrc/einsum.cu), which addresses all operands through
per-label strides, so none of the special cases above needs a special kernel.  ``einsum`` is an
API-completeness row, not a hot path: large matrix products belong to :func:`mygrad_b200.matmul`, which
runs on the convolution GEMM kernels.

Divergences: a single-operand einsum without summation returns a copy, not a view (device arrays carry
no strides); ``out=`` is not supported; ``optimize`` is accepted and ignored (there is no pairwise
contraction path to choose).
"""
from __future__ import annotations

import ctypes as C
import string
from collections import Counter
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import check, lib
from .device_array import DeviceArray, asdevice, current_stream, zeros
from .errors import SkipGradient
from .operation_base import Operation
from .tensor_base import Tensor

__all__ = ["einsum", "EinSum"]

# internal names of the dimensions an ellipsis stands for: never letters, so they cannot collide with user labels
_ELLIPSIS_POOL = "0123456789!@#$%^&*"


def _i64(values) -> C.Array:
    values = [int(v) for v in values]
    return (C.c_int64 * max(1, len(values)))(*values)


def _strides(shape: Sequence[int]) -> List[int]:
    out, acc = [], 1
    for s in reversed(shape):
        out.append(acc)
        acc *= s
    return out[::-1]


def _label_strides(labels: str, shape: Sequence[int], wanted: str) -> List[int]:
    """element stride of a C-contiguous array of `shape` (axes named by `labels`) per label in `wanted`:
    repeated labels add up (diagonal), extent-1 axes and absent labels give 0 (broadcast)"""
    st = _strides(shape)
    return [sum(s for lb, s, n in zip(labels, st, shape) if lb == w and n != 1) for w in wanted]


def _contract(dtype, operands: Sequence[DeviceArray], op_labels: Sequence[str], sizes: Dict[str, int],
              out_labels: str, out: DeviceArray, out_full_labels: str) -> None:
    """out[...] = sum over the labels not in `out_labels` of the product of the operands.  `out` is an array
    whose axes are named by `out_full_labels` (which may repeat labels: the unique ones are `out_labels`)."""
    summed = "".join(sorted(set("".join(op_labels)) - set(out_labels)))
    nout, nsum, nops = len(out_labels), len(summed), len(operands)
    ptrs = (C.c_void_p * nops)(*[C.c_void_p(o.ptr) for o in operands])
    oos = [s for lb, op in zip(op_labels, operands) for s in _label_strides(lb, op.shape, out_labels)]
    oss = [s for lb, op in zip(op_labels, operands) for s in _label_strides(lb, op.shape, summed)]
    out_stride = _label_strides(out_full_labels, [sizes[lb] for lb in out_full_labels], out_labels)
    # an output label of extent 1 has stride 0 above; harmless (its only index is 0)
    check(lib.mgb_einsum(0 if np.dtype(dtype) == np.float32 else 1, nops, ptrs, nout,
                         _i64(sizes[lb] for lb in out_labels), _i64(out_stride), _i64(oos), nsum,
                         _i64(sizes[lb] for lb in summed), _i64(oss), C.c_void_p(out.ptr),
                         C.c_void_p(current_stream())))


def _unique_keep_last(lbl: str) -> str:
    """'ijikik' -> 'jik': drop all but the LAST occurrence of every label (ops.py:22-43)"""
    seen, out = set(), []
    for ch in reversed(lbl):
        if ch not in seen:
            seen.add(ch)
            out.append(ch)
    return "".join(reversed(out))


class EinSum(Operation):
    can_return_view = False     # the reference may return a view (ops.py:80); device arrays cannot

    def __call__(self, *variables: Tensor, in_lbls: str, out_lbls: str, optimize=False) -> DeviceArray:
        self.in_lbls = in_lbls.split(",")
        self.out_lbls = out_lbls
        self.variables = variables
        self.optimize = optimize
        self._cache: Optional[Counter] = None
        dtype = np.result_type(*[v.data.dtype for v in variables])
        self._dtype = dtype
        self._arrays = [v.data.astype(dtype, copy=False) for v in variables]
        self._sizes = _sizes(self.in_lbls, [a.shape for a in self._arrays])
        out = DeviceArray(tuple(self._sizes[lb] for lb in out_lbls), dtype)
        _contract(dtype, self._arrays, self.in_lbls, self._sizes, out_lbls, out, out_lbls)
        return out

    @property
    def cache(self) -> Counter:
        # how often each (tensor, subscript) pair feeds the einsum: its gradient is computed once and scaled
        if self._cache is None:
            self._cache = Counter(zip((id(v) for v in self.variables), self.in_lbls))
        return self._cache

    def backward_var(self, grad: DeviceArray, index: int, **kwargs) -> DeviceArray:
        """fwd "ijk,k->ji" (x, y):  dx = "ji,k->ijk" (grad, y),  dy = "ji,ijk->k" (grad, x)   (ops.py:130-243)"""
        in_lbls = list(self.in_lbls)
        var_lbl_full = in_lbls.pop(index)
        var = self.variables[index]
        factor = self.cache[(id(var), var_lbl_full)]
        if factor == 0:
            raise SkipGradient()        # this tensor-subscript pair has already been back-propagated
        self.cache[(id(var), var_lbl_full)] = 0

        var_lbl = _unique_keep_last(var_lbl_full)
        repeats = len(var_lbl) != len(var_lbl_full)
        operands = [grad.astype(self._dtype, copy=False)] + self._arrays[:index] + self._arrays[index + 1:]
        op_labels = [self.out_lbls] + in_lbls
        # the gradient at the full extent of every label; axes of extent 1 in `var` that were broadcast
        # (ellipsis dimensions) are summed away afterwards, exactly like reduce_broadcast in ops.py:203
        full_shape = tuple(self._sizes[lb] for lb in var_lbl_full)
        dfdx = zeros(full_shape, self._dtype) if repeats else DeviceArray(full_shape, self._dtype)
        # labels of `var` that no other operand and not the output carries were summed without a partner:
        # every operand has stride 0 along them, i.e. the gradient is broadcast (ops.py:168-190)
        _contract(self._dtype, operands, op_labels, self._sizes, var_lbl, dfdx, var_lbl_full)
        if full_shape != tuple(var.shape):
            dfdx = dfdx.sum_to_shape(var.shape)
        if factor > 1:
            dfdx *= float(factor)
        return dfdx


def _sizes(in_lbls: Sequence[str], shapes: Sequence[Tuple[int, ...]]) -> Dict[str, int]:
    """label -> extent.  Named labels must agree exactly; ellipsis dimensions broadcast (1 or equal)."""
    sizes: Dict[str, int] = {}
    for k, (lbl, shape) in enumerate(zip(in_lbls, shapes)):
        if len(lbl) != len(shape):
            raise ValueError(f"einstein sum subscripts string contains too many or too few subscripts for operand {k}")
        for ch, n in zip(lbl, shape):
            old = sizes.get(ch)
            if old is None or old == n:
                sizes[ch] = n
            elif ch in _ELLIPSIS_POOL and (old == 1 or n == 1):
                sizes[ch] = max(old, n)
            elif ch in _ELLIPSIS_POOL:
                raise ValueError(f"operands could not be broadcast together: ellipsis dimension {old} vs {n}")
            else:
                raise ValueError(f"operands could not be broadcast together with remapped shapes: dimension "
                                 f"'{ch}' is {n} for operand {k} but {old} elsewhere")
    return sizes


def _parse(operands) -> Tuple[List[str], str, list]:
    """subscripts (string or sublist form) -> per-operand label strings with the ellipsis expanded into
    internal labels, the output labels, and the operands as Tensors  (numpy.einsum's rules)"""
    if len(operands) == 0:
        raise ValueError("No input operands")
    if isinstance(operands[0], str):
        subs = operands[0].replace(" ", "")
        ops = list(operands[1:])
        for ch in subs:
            if ch not in string.ascii_letters + ".,->":
                raise ValueError(f"Character {ch} is not a valid symbol.")
    else:   # einsum(op0, sublist0, op1, sublist1, ..., [sublistout])
        rest = list(operands)
        ops, parts = [], []
        while len(rest) >= 2:
            ops.append(rest.pop(0))
            parts.append(rest.pop(0))

        def conv(sub) -> str:
            try:
                return "".join("..." if item is Ellipsis else string.ascii_letters[int(item)] for item in sub)
            except (TypeError, IndexError, ValueError) as e:
                raise TypeError("For this input type lists must contain either int or Ellipsis") from e

        subs = ",".join(conv(p) for p in parts)
        if rest:
            subs += "->" + conv(rest[0])
    if ("-" in subs or ">" in subs) and (subs.count("->") != 1 or subs.count("-") != 1 or subs.count(">") != 1):
        raise ValueError("Subscripts can only contain one '->'.")
    explicit = "->" in subs
    in_part, out_part = subs.split("->") if explicit else (subs, None)
    in_terms = in_part.split(",")
    if len(in_terms) != len(ops):
        raise ValueError("Number of einsum subscripts must be equal to the number of operands.")
    arrays = [op if isinstance(op, Tensor) else Tensor(asdevice(op), copy=False, constant=True) for op in ops]

    used = set(ch for ch in subs if ch in string.ascii_letters)
    pool = list(_ELLIPSIS_POOL)
    longest, expanded = 0, []
    for term, arr in zip(in_terms, arrays):
        if "." in term:
            if term.count(".") != 3 or term.count("...") != 1:
                raise ValueError("Invalid Ellipses.")
            n_ell = arr.ndim - (len(term) - 3)
            if n_ell < 0:
                raise ValueError("Ellipses lengths do not match.")
            if n_ell > len(pool):
                raise ValueError("too many subscripts in einsum")
            longest = max(longest, n_ell)
            # right-aligned: the LAST ellipsis dimension of every operand carries the same label
            term = term.replace("...", "".join(pool[:n_ell][::-1]))
        expanded.append(term)
    ell_out = "".join(pool[:longest][::-1])
    joined = "".join(expanded)
    if explicit:
        if "." in out_part and (out_part.count(".") != 3 or out_part.count("...") != 1):
            raise ValueError("Invalid Ellipses.")
        out_lbls = out_part.replace("...", ell_out)
        for ch in out_lbls:
            if out_lbls.count(ch) != 1:
                raise ValueError(f"einstein sum subscripts string includes output subscript '{ch}' multiple times")
            if ch not in joined:
                raise ValueError(f"einstein sum subscripts string included output subscript '{ch}' which never "
                                 "appeared in an input")
        # an input-side ellipsis that the explicit output omits is summed away — the reference's parser
        # accepts "...,...->" and "...i->i" (linalg/funcs.py), unlike numpy's C-level einsum
    else:
        # implicit mode: ellipsis dimensions first, then the labels that appear exactly once, alphabetically
        out_lbls = ell_out + "".join(ch for ch in sorted(used) if joined.count(ch) == 1)
    return expanded, out_lbls, arrays


def einsum(*operands, optimize=False, out=None, constant: Optional[bool] = None) -> Tensor:
    """``einsum(subscripts, *operands)`` — the Einstein summation convention on device tensors, with
    back-propagation through every operand (``mygrad.einsum``, linalg/funcs.py:178-440)."""
    if out is not None:
        raise NotImplementedError("einsum(out=...) is not supported on device tensors")
    in_lbls, out_lbls, arrays = _parse(operands)
    return Tensor._op(EinSum, *arrays, op_kwargs=dict(in_lbls=",".join(in_lbls), out_lbls=out_lbls,
                                                     optimize=optimize), constant=constant)
