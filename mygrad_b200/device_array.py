"""DeviceArray: a C-contiguous float32/float64 buffer in B200 HBM.

It plays the role ``numpy.ndarray`` plays as ``Tensor.data`` in the reference and
provides exactly the attributes the graph runtime touches on it:
``shape, ndim, dtype, size, base, astype, copy, +=, full_like`` plus host transfer.
Buffers are exchanged with other libraries through ``__cuda_array_interface__``
(version 3); memory comes from the stream-ordered pool behind ``mgb_malloc_async``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Optional, Sequence, Tuple, Union

import numpy as np

from ._lib import F32, F64, check, lib

__all__ = [
    "DeviceArray",
    "asdevice",
    "empty",
    "zeros",
    "full",
    "full_like",
    "pinned_empty",
    "current_stream",
    "set_stream",
    "synchronize",
    "device_available",
    "empty_cache",
]

_SUPPORTED = (np.dtype(np.float32), np.dtype(np.float64))
_stream: int = 0  # legacy default stream: ordered with torch's default stream


def current_stream() -> int:
    return _stream


def set_stream(stream: Optional[int]) -> None:
    """Routes all subsequent launches to `stream` (a ``cudaStream_t`` as int; None/0 = default)."""
    global _stream
    _stream = int(stream or 0)


def synchronize() -> None:
    check(lib.mgb_stream_synchronize(C.c_void_p(_stream)))


def device_available() -> bool:
    n = C.c_int(0)
    return lib.mgb_device_count(C.byref(n)) == 0 and n.value > 0


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError(f"mygrad_b200 supports float32 and float64 device arrays, got {dt}")


# Exact-size block cache in front of the stream-ordered pool.  A training loop asks for
# the same sizes every step; handing a freed block straight to the next request of that
# size costs no driver call and never makes the pool grow or trim (which shows up as
# multi-millisecond hiccups with GB-sized activations).  Re-use is safe in stream order
# ONLY for blocks whose whole life was spent on the stream they were allocated on: every
# kernel of the previous owner was then enqueued on that stream before the new owner's.
# A block that was (or may have been) used on another stream — the current stream changed
# while it was alive, or ``record_stream`` named a side stream — is not cached: it goes back
# to the driver pool behind events that order the free after the work on those streams.
_CACHE: dict = {}          # (device, stream, nbytes) -> [ptr, ...]
_GRAPH_STREAMS: set = set()   # streams owned by a live StepGraph: their cached blocks are part of the recorded graph
_capturing: bool = False      # inside StepGraph capture: a cache miss would put a graph-owned allocation into the graph
_CACHE_LIMIT = int(float(os.environ.get("MGB_CACHE_LIMIT", 32e9)))   # bytes kept at most
_cached_bytes = 0


def _device() -> int:
    d = C.c_int(0)
    return d.value if lib.mgb_get_device(C.byref(d)) == 0 else 0


def empty_cache() -> None:
    """returns every cached block to the driver's stream-ordered pool"""
    global _cached_bytes
    cur = _device()
    for key, ptrs in list(_CACHE.items()):
        dev, stream, nbytes = key
        if stream in _GRAPH_STREAMS:      # a recorded graph replays into these blocks: they stay
            continue
        if dev != cur:
            lib.mgb_set_device(dev)
        for p in ptrs:
            lib.mgb_free_async(C.c_void_p(p), C.c_void_p(stream))
        if dev != cur:
            lib.mgb_set_device(cur)
        _cached_bytes -= nbytes * len(ptrs)
        del _CACHE[key]


def _order_after(stream: int, others) -> None:
    """makes `stream` wait for everything queued so far on each stream of `others`"""
    for o in others:
        if o == stream:
            continue
        e = C.c_void_p()
        if lib.mgb_event_create(C.byref(e)) != 0:
            lib.mgb_stream_synchronize(C.c_void_p(o))      # no event to be had: fall back to a host wait
            continue
        lib.mgb_event_record(e, C.c_void_p(o))
        lib.mgb_stream_wait_event(C.c_void_p(stream), e)
        lib.mgb_event_destroy(e)


class _Allocation:
    """owns one device allocation; recycled (stream-ordered) when the last view dies"""

    __slots__ = ("ptr", "nbytes", "stream", "device", "foreign")

    def __init__(self, nbytes: int):
        global _cached_bytes
        self.nbytes = max(16, (nbytes + 511) // 512 * 512)
        self.stream = _stream
        self.device = _device()
        self.foreign = None          # set of other streams this block was used on
        free = _CACHE.get((self.device, self.stream, self.nbytes))
        if free:
            self.ptr = free.pop()
            _cached_bytes -= self.nbytes
            return
        if _capturing:
            raise RuntimeError(
                f"StepGraph: the captured step asked for a {self.nbytes}-byte block the warm-up steps never freed "
                "(its allocations differ from step to step); graphs need a step whose buffer sizes repeat")
        p = C.c_void_p()
        rc = lib.mgb_malloc_async(C.byref(p), C.c_size_t(self.nbytes), C.c_void_p(_stream))
        if rc != 0 and _cached_bytes:
            # the driver pool is exhausted while blocks sit idle in the cache: hand them back and retry once
            empty_cache()
            lib.mgb_device_synchronize()
            rc = lib.mgb_malloc_async(C.byref(p), C.c_size_t(self.nbytes), C.c_void_p(_stream))
        check(rc)
        self.ptr = int(p.value or 0)

    def __del__(self):
        try:
            global _cached_bytes
            if not self.ptr:
                return
            others = set(self.foreign or ())
            if _stream != self.stream:
                others.add(_stream)          # alive across a set_stream(): assume it was used there too
            if not others and _cached_bytes + self.nbytes <= _CACHE_LIMIT and self.device == _device():
                _CACHE.setdefault((self.device, self.stream, self.nbytes), []).append(self.ptr)
                _cached_bytes += self.nbytes
            else:
                _order_after(self.stream, others)
                lib.mgb_free_async(C.c_void_p(self.ptr), C.c_void_p(self.stream))
            self.ptr = 0
        except Exception:  # interpreter shutdown
            pass


class DeviceArray:
    __slots__ = ("_alloc", "_keepalive", "ptr", "shape", "dtype", "base", "reduced", "__weakref__")

    def __init__(self, shape: Sequence[int], dtype=np.float32, *, _alloc=None, _ptr=None,
                 _keepalive=None, _base=None):
        if isinstance(shape, (int, np.integer)):
            shape = (shape,)
        self.shape: Tuple[int, ...] = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        if self.dtype not in _SUPPORTED:
            raise TypeError(f"mygrad_b200 supports float32 and float64 device arrays, got {self.dtype}")
        self.base = _base
        self._keepalive = _keepalive
        # True when the buffer already holds the SUM over all data-parallel ranks (fused wgrad,
        # cross-rank batchnorm): DataParallel.allreduce then only scales it (dist.py)
        self.reduced = bool(getattr(_base, "reduced", False)) if _base is not None else False
        if _ptr is not None:
            self._alloc = _alloc
            self.ptr = int(_ptr)
        else:
            self._alloc = _Allocation(self.nbytes)
            self.ptr = self._alloc.ptr

    # ---- ndarray-like metadata -------------------------------------------------
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def itemsize(self) -> int:
        return self.dtype.itemsize

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def code(self) -> int:
        return F32 if self.dtype == np.float32 else F64

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    @property
    def __cuda_array_interface__(self) -> dict:
        return {
            "shape": self.shape,
            "typestr": self.dtype.str,
            "data": (self.ptr if self.size else 0, False),
            "strides": None,
            "version": 3,
            "stream": _stream if _stream else 1,  # 1 = legacy default stream in CAI v3
        }

    def record_stream(self, stream) -> None:
        """declares that work on `stream` (not the allocating stream) uses this buffer: its memory is
        then released behind an event on that stream instead of being recycled in allocation-stream order"""
        s = int(getattr(stream, "value", stream) or 0)
        a = self._alloc if self._alloc is not None else getattr(self.base, "_alloc", None)
        if a is not None and s != a.stream:
            if a.foreign is None:
                a.foreign = set()
            a.foreign.add(s)

    # ---- construction / transfer -------------------------------------------------
    @classmethod
    def from_numpy(cls, arr: np.ndarray) -> "DeviceArray":
        if not arr.flags["C_CONTIGUOUS"]:
            arr = np.ascontiguousarray(arr)   # (0-d arrays are contiguous; ascontiguousarray would make them 1-d)
        out = cls(arr.shape, arr.dtype)
        if arr.size:
            check(lib.mgb_memcpy_h2d(C.c_void_p(out.ptr), C.c_void_p(arr.ctypes.data),
                                     C.c_size_t(arr.nbytes), C.c_void_p(_stream)))
            # pageable sources are staged before the call returns; pinned ones are not
            check(lib.mgb_stream_synchronize(C.c_void_p(_stream)))
        return out

    def set_from_host(self, arr: np.ndarray, *, sync: bool = True) -> None:
        """H2D copy into this buffer (arr must match shape/dtype; pinned arrays copy async)."""
        if tuple(arr.shape) != self.shape or arr.dtype != self.dtype or not arr.flags["C_CONTIGUOUS"]:
            raise ValueError("set_from_host: host array must be C-contiguous with matching shape/dtype")
        if arr.size:
            check(lib.mgb_memcpy_h2d(C.c_void_p(self.ptr), C.c_void_p(arr.ctypes.data),
                                     C.c_size_t(arr.nbytes), C.c_void_p(_stream)))
            if sync:
                check(lib.mgb_stream_synchronize(C.c_void_p(_stream)))

    def get(self, out: Optional[np.ndarray] = None, *, sync: bool = True) -> np.ndarray:
        """D2H copy; returns a NumPy array (writes into `out` if given)."""
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        elif tuple(out.shape) != self.shape or out.dtype != self.dtype or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("get: `out` must be C-contiguous with matching shape/dtype")
        if self.size:
            check(lib.mgb_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(self.ptr),
                                     C.c_size_t(self.nbytes), C.c_void_p(_stream)))
            if sync:
                check(lib.mgb_stream_synchronize(C.c_void_p(_stream)))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None and np.dtype(dtype) != a.dtype else a

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(()).item()

    # ---- the operations the graph runtime performs on grads ------------------------
    def copy(self) -> "DeviceArray":
        out = DeviceArray(self.shape, self.dtype)
        check(lib.mgb_copy_cast(self.code, self.code, C.c_size_t(self.size), C.c_void_p(self.ptr),
                                C.c_void_p(out.ptr), C.c_void_p(_stream)))
        out.reduced = self.reduced
        return out

    def astype(self, dtype, copy: bool = True) -> "DeviceArray":
        dt = np.dtype(dtype)
        if dt == self.dtype and not copy:
            return self
        out = DeviceArray(self.shape, dt)
        check(lib.mgb_copy_cast(self.code, out.code, C.c_size_t(self.size), C.c_void_p(self.ptr),
                                C.c_void_p(out.ptr), C.c_void_p(_stream)))
        out.reduced = self.reduced
        return out

    def __iadd__(self, other: "DeviceArray") -> "DeviceArray":
        other = asdevice(other)
        if other.shape != self.shape:
            raise ValueError(f"in-place add needs equal shapes, got {self.shape} and {other.shape}")
        if other.dtype != self.dtype:
            other = other.astype(self.dtype)
        if self.reduced != other.reduced:
            raise RuntimeError(
                "accumulating a gradient that is already summed over the data-parallel ranks with a rank-local "
                "one: all-reduce the local contribution first (DataParallel.allreduce) or keep fused_wgrad() / "
                "dp= off for parameters that also receive rank-local gradients")
        check(lib.mgb_axpy(self.code, C.c_size_t(self.size), C.c_void_p(other.ptr),
                           C.c_void_p(self.ptr), C.c_void_p(_stream)))
        return self

    def __imul__(self, alpha: float) -> "DeviceArray":
        check(lib.mgb_scale(self.code, C.c_size_t(self.size), C.c_double(float(alpha)),
                            C.c_void_p(self.ptr), C.c_void_p(_stream)))
        return self

    def fill(self, value: float) -> "DeviceArray":
        check(lib.mgb_fill(self.code, C.c_size_t(self.size), C.c_double(float(value)),
                           C.c_void_p(self.ptr), C.c_void_p(_stream)))
        return self

    def reshape(self, *shape) -> "DeviceArray":
        """metadata-only view (buffers are always C-contiguous)"""
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(int(s) for s in shape)
        if shape.count(-1) > 1:
            raise ValueError("can only specify one unknown dimension")
        if -1 in shape:
            known = 1
            for s in shape:
                if s != -1:
                    known *= s
            if known == 0 or self.size % known:
                raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
            shape[shape.index(-1)] = self.size // known
        n = 1
        for s in shape:
            n *= s
        if n != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
        root = self.base if self.base is not None else self
        return DeviceArray(shape, self.dtype, _alloc=self._alloc, _ptr=self.ptr,
                           _keepalive=self._keepalive, _base=root)

    def sum_to_shape(self, var_shape: Sequence[int]) -> "DeviceArray":
        """un-broadcast reduction to `var_shape` (reduce_broadcast, _utils/__init__.py:131-168)"""
        var_shape = tuple(int(s) for s in var_shape)
        out = DeviceArray(var_shape, self.dtype)
        gs = (C.c_int64 * max(1, self.ndim))(*self.shape)
        vs = (C.c_int64 * max(1, len(var_shape)))(*var_shape)
        check(lib.mgb_reduce_broadcast(self.code, self.ndim, gs, len(var_shape), vs,
                                       C.c_void_p(self.ptr), C.c_void_p(out.ptr), 0,
                                       C.c_void_p(_stream)))
        out.reduced = self.reduced
        return out

    def __repr__(self) -> str:
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype.name}, ptr=0x{self.ptr:x})"


ArrayLike = Union[DeviceArray, np.ndarray, Any]


def asdevice(obj: ArrayLike, dtype=None) -> DeviceArray:
    """DeviceArray from a DeviceArray (no copy), any ``__cuda_array_interface__`` producer
    (zero-copy, must be C-contiguous f32/f64), or host data (uploaded)."""
    if isinstance(obj, DeviceArray):
        out = obj
    elif hasattr(obj, "__cuda_array_interface__"):
        cai = obj.__cuda_array_interface__
        if cai.get("strides") is not None:
            # accept explicitly given C-contiguous strides
            shape, item = cai["shape"], np.dtype(cai["typestr"]).itemsize
            expect, acc = [], item
            for s in reversed(shape):
                expect.append(acc)
                acc *= s
            if tuple(cai["strides"]) != tuple(reversed(expect)):
                raise ValueError("device arrays must be C-contiguous")
        # CAI v3: 1 = legacy default stream, 2 = per-thread default, else a cudaStream_t
        producer_stream = cai.get("stream")
        if producer_stream is not None and int(producer_stream) != (_stream or 1):
            check(lib.mgb_stream_synchronize(C.c_void_p(int(producer_stream))))
        out = DeviceArray(cai["shape"], np.dtype(cai["typestr"]), _ptr=cai["data"][0] or 0,
                          _keepalive=obj)
    else:
        arr = np.asarray(obj)
        if arr.dtype not in _SUPPORTED:
            if dtype is None:
                # mirror NumPy defaults: python floats / ints become float64
                arr = arr.astype(np.float64)
            else:
                arr = arr.astype(dtype)
        out = DeviceArray.from_numpy(arr)
    if dtype is not None and np.dtype(dtype) != out.dtype:
        out = out.astype(dtype)
    return out


def empty(shape, dtype=np.float32) -> DeviceArray:
    return DeviceArray(shape if isinstance(shape, (tuple, list)) else (shape,), dtype)


def full(shape, value: float, dtype=np.float32) -> DeviceArray:
    return empty(shape, dtype).fill(value)


def zeros(shape, dtype=np.float32) -> DeviceArray:
    out = empty(shape, dtype)
    check(lib.mgb_memset(C.c_void_p(out.ptr), 0, C.c_size_t(out.nbytes), C.c_void_p(_stream)))
    return out


def full_like(a: DeviceArray, value: float) -> DeviceArray:
    return DeviceArray(a.shape, a.dtype).fill(value)


class _PinnedBlock:
    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(lib.mgb_host_alloc(C.byref(p), C.c_size_t(nbytes)))
        self.ptr = int(p.value)

    def __del__(self):
        try:
            lib.mgb_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """NumPy array over page-locked host memory (async H2D/D2H source/target)."""
    dt = np.dtype(dtype)
    shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
    n = int(np.prod(shape)) if shape else 1
    block = _PinnedBlock(max(1, n) * dt.itemsize)
    buf = (C.c_char * (max(1, n) * dt.itemsize)).from_address(block.ptr)
    buf._block = block  # np.frombuffer keeps `buf` alive, `buf` keeps the pinned block alive
    return np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
