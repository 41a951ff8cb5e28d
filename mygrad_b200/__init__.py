"""mygrad_b200 — B200-native (sm_100a) hot path for MyGrad's conv_nd / max_pool.

The public surface mirrors the part of ``mygrad`` that the path touches:

    from mygrad_b200 import Tensor, Operation, no_autodiff, execute_op
    from mygrad_b200.nnet.layers import conv_nd, max_pool

Arithmetic runs in hand-written CUDA kernels behind a C ABI (include/mgb.h), called
through ctypes; tensors live in HBM as ``DeviceArray`` and are exchanged through
``__cuda_array_interface__``.  There is no CPU fallback: importing this package
without the built extension raises ImportError.
"""
from . import _graph_tracking
from ._graph_tracking import no_autodiff
from ._lib import LIB_PATH, lib
from .device_array import (
    DeviceArray,
    asdevice,
    current_stream,
    device_available,
    empty,
    empty_cache,
    full,
    full_like,
    pinned_empty,
    set_stream,
    synchronize,
    zeros,
)
from .errors import DisconnectedView, InvalidBackprop, InvalidGradient, SkipGradient
from .operation_base import Operation
from .tensor_base import Tensor, asarray, tensor
from . import nnet
from .host import HostLayerPipeline
from .nnet.layers import conv_nd, conv_pool, max_pool, mean_pool
from .math_ops import add, matmul, reshape
from ._io import load, save
from .linalg import einsum
from .graph import StepGraph

execute_op = Tensor._op  # same alias as mygrad.execute_op (src/mygrad/__init__.py:52)

__version__ = "0.1.0"

__all__ = [
    "Tensor", "tensor", "asarray", "Operation", "execute_op", "no_autodiff",
    "DeviceArray", "asdevice", "empty", "zeros", "full", "full_like", "pinned_empty",
    "current_stream", "set_stream", "synchronize", "device_available", "empty_cache",
    "InvalidGradient", "InvalidBackprop", "DisconnectedView", "SkipGradient",
    "nnet", "conv_nd", "max_pool", "mean_pool", "conv_pool", "add", "matmul", "reshape", "einsum", "save", "load", "HostLayerPipeline", "StepGraph",
    "lib", "LIB_PATH",
]
