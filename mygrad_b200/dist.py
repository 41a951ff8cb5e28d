"""Batch-axis data parallelism: one process per GPU, dW / bias grads summed by NCCL.

The reference has no notion of ranks.  Every output sample of conv
fwd, dX, pool fwd and pool bwd depends only on the same input sample, so rank r of R
works on ``x[lo:hi]`` with a full replica of the filters; only parameter gradients
(dW, broadcast biases) contract over the batch and need one sum all-reduce.  The
result to match is the single-process reference on the full batch.

``torch.distributed`` is used for rendezvous only (shipping the NCCL unique id, the CUDA IPC
handles and barriers).  Two collectives sit underneath:

* ``mgb_peer_allreduce_sum`` / ``mgb_conv_wgrad_dp`` — our own one-kernel all-reduce over NVLink
  peer memory (csrc/comm.cu): parameter gradients are a few hundred KB, i.e. pure latency, and the
  dW one always follows the split-K reduction of the wgrad GEMM, so both run in ONE kernel
  (``with dp.fused_wgrad(): loss.backward()``).  Results are summed in rank order in float64:
  bit-identical on every rank.
* ``mgb_allreduce_sum`` — ncclAllReduce on a side stream, ordered against the compute stream with
  events, for buffers larger than a peer slot or when CUDA IPC is unavailable (``MGB_DP_PEER=0``
  forces it).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import check, lib
from .device_array import DeviceArray, current_stream

__all__ = ["shard_bounds", "scale_for_mean_loss", "DataParallel", "fused_dp"]

_PEER_SLOT_BYTES = 8 << 20     # per-call capacity of the peer-memory all-reduce (dW of config 3 is 6.55 MB)
# the one-shot kernel reads R-1 copies of the message: it wins while the message is latency-sized
# (295 KB at 8 GPUs: 18 us vs 28 us for NCCL) and loses above ~1 MB (6.5 MB: 205 vs 65 us); MGB_DP_PEER_MAX overrides
_PEER_MAX_BYTES = int(os.environ.get("MGB_DP_PEER_MAX", 1 << 20))
_FUSED: Optional["DataParallel"] = None


def fused_dp() -> Optional["DataParallel"]:
    """the DataParallel whose ``fused_wgrad()`` context is active (ConvND.backward_var asks)"""
    return _FUSED


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous batch slice [lo, hi) of rank `rank`; the first n % world ranks get one extra"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scale_for_mean_loss(local_n: int, global_n: int) -> float:
    """factor for a loss that averages over the LOCAL batch (e.g. softmax_crossentropy,
    nnet/losses/softmax_crossentropy.py:41,46) so that the summed all-reduce equals the
    full-batch mean gradient"""
    return float(local_n) / float(global_n)


class DataParallel:
    """Owns the per-process NCCL communicator and a communication stream."""

    def __init__(self, rank: Optional[int] = None, world_size: Optional[int] = None,
                 local_rank: Optional[int] = None):
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
        self.local_rank = int(os.environ.get("LOCAL_RANK", self.rank)) if local_rank is None else local_rank
        self._comm_stream = None
        self._event = None
        self._ready = False
        self._buckets: dict = {}
        self.peer_slot_bytes = 0          # > 0 once the peer-memory group is connected
        self.peer_error: Optional[str] = None

    # -- setup ---------------------------------------------------------------------
    def init(self) -> "DataParallel":
        """binds this process to GPU `local_rank` and creates the communicator; with
        world_size > 1 a torch.distributed process group must already be initialised"""
        check(lib.mgb_set_device(self.local_rank))
        s, e = C.c_void_p(), C.c_void_p()
        check(lib.mgb_stream_create(C.byref(s)))
        check(lib.mgb_event_create(C.byref(e)))
        self._comm_stream, self._event = s, e
        if self.world_size > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("torch.distributed must be initialised before DataParallel.init()")
            uid = (C.c_char * 128)()
            if self.rank == 0:
                check(lib.mgb_comm_unique_id(uid))
            box = [bytes(uid.raw) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            buf = (C.c_char * 128).from_buffer_copy(box[0])
            check(lib.mgb_comm_init_rank(buf, self.world_size, self.rank))
            if os.environ.get("MGB_DP_PEER", "1") != "0":
                self._init_peer(dist)
        self._ready = True
        return self

    def _init_peer(self, dist) -> None:
        """maps every rank's staging region into this process (CUDA IPC over NVLink); on any failure
        the NCCL path stays in charge — all ranks decide together"""
        handle = (C.c_char * 64)()
        ok = lib.mgb_peer_create(self.rank, self.world_size, C.c_size_t(_PEER_SLOT_BYTES), handle) == 0
        err = None if ok else lib.mgb_last_error().decode("utf-8", "replace")
        boxes = [None] * self.world_size
        dist.all_gather_object(boxes, (ok, bytes(handle.raw), err))
        if all(b[0] for b in boxes):
            blob = b"".join(b[1] for b in boxes)
            ok = lib.mgb_peer_connect((C.c_char * len(blob)).from_buffer_copy(blob)) == 0
            err = None if ok else lib.mgb_last_error().decode("utf-8", "replace")
        else:
            ok, err = False, next(b[2] for b in boxes if not b[0])
        flags = [None] * self.world_size
        dist.all_gather_object(flags, ok)
        if all(flags):
            n = C.c_size_t(0)
            check(lib.mgb_peer_slot_bytes(C.byref(n)))
            self.peer_slot_bytes = int(n.value)
        else:
            lib.mgb_peer_destroy()
            self.peer_error = err or "a peer rank could not map the staging regions"

    def fuses(self, nbytes: int) -> bool:
        """True when a gradient of `nbytes` goes through the peer-memory kernels (fused wgrad, bucket
        all-reduce) rather than NCCL"""
        return (self.world_size > 1 and 0 < nbytes <= min(self.peer_slot_bytes, _PEER_MAX_BYTES))

    def fused_wgrad(self):
        """context manager: conv filter gradients produced inside it are ALREADY summed over the ranks
        (mgb_conv_wgrad_dp: split-K reduction and cross-rank sum in one kernel).  Such buffers carry
        ``DeviceArray.reduced = True`` (so do dgamma / dbeta of ``batchnorm(dp=...)``), and
        ``allreduce`` only scales them — the caller passes every parameter gradient and never has to
        re-derive which path a gradient took.  Filters larger than the peer kernel's sweet spot
        (``fuses(nbytes)`` is False) keep the plain wgrad, stay rank-local and are all-reduced as usual."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            global _FUSED
            prev = _FUSED
            on = self.world_size > 1 and self.peer_slot_bytes > 0 and os.environ.get("MGB_DP_FUSED", "1") != "0"
            _FUSED = self if on else None
            try:
                yield self
            finally:
                _FUSED = prev

        return ctx()

    def close(self) -> None:
        if self.peer_slot_bytes:
            lib.mgb_peer_destroy()
            self.peer_slot_bytes = 0
        if self._ready and self.world_size > 1:
            lib.mgb_comm_destroy()
        if self._comm_stream is not None:
            lib.mgb_stream_destroy(self._comm_stream)
            lib.mgb_event_destroy(self._event)
            self._comm_stream = self._event = None
        self._ready = False
        self._buckets.clear()

    # -- sharding --------------------------------------------------------------------
    def shard(self, n: int) -> Tuple[int, int]:
        return shard_bounds(n, self.rank, self.world_size)

    # -- gradient exchange -------------------------------------------------------------
    def allreduce_async(self, grads: Iterable[DeviceArray]) -> None:
        """sum all-reduce of each buffer on the communication stream, after everything
        already queued on the compute stream (later compute work overlaps with it)"""
        if self.world_size == 1:
            return
        grads = list(grads)
        compute = C.c_void_p(current_stream())
        check(lib.mgb_event_record(self._event, compute))
        check(lib.mgb_stream_wait_event(self._comm_stream, self._event))
        for g in grads:
            check(lib.mgb_allreduce_sum(C.c_void_p(g.ptr), C.c_size_t(g.size), g.code, self._comm_stream))

    def wait(self) -> None:
        """makes the compute stream wait for the outstanding all-reduces"""
        if self.world_size == 1:
            return
        check(lib.mgb_event_record(self._event, self._comm_stream))
        check(lib.mgb_stream_wait_event(C.c_void_p(current_stream()), self._event))

    def _reduce_flat(self, flat: DeviceArray) -> None:
        """one collective on one contiguous buffer: the peer-memory kernel when it fits, else NCCL"""
        if self.fuses(flat.nbytes):
            check(lib.mgb_peer_allreduce_sum(C.c_void_p(flat.ptr), C.c_size_t(flat.size), flat.code,
                                             C.c_void_p(current_stream())))
        else:
            self.allreduce_async([flat])
            self.wait()

    def allreduce(self, grads: Sequence[DeviceArray], scale: Optional[float] = None,
                  summed: Sequence[DeviceArray] = ()) -> None:
        """sum all-reduce (then optional scaling) of every buffer.  Parameter gradients are small
        (C5: 0.26 MB in six tensors), so the cost is per-call latency: buffers of one dtype are
        packed into a single bucket, reduced with ONE collective and unpacked.  Buffers whose
        ``reduced`` flag is set (fused wgrad, cross-rank batchnorm) — or that are listed in ``summed`` —
        already hold the sum over the ranks and are only scaled."""
        grads = list(grads)
        extra = {id(g) for g in summed}
        pre = [g for g in grads if g.reduced or id(g) in extra]
        pre += [g for g in summed if not any(g is h for h in grads)]
        if scale is not None and scale != 1.0:
            for g in pre:
                g *= scale
        grads = [g for g in grads if not (g.reduced or id(g) in extra)]
        if self.world_size == 1:
            if scale is not None and scale != 1.0:
                for g in grads:
                    g *= scale
            return
        by_dtype: dict = {}
        for g in grads:
            by_dtype.setdefault(g.dtype.str, []).append(g)
        for group in by_dtype.values():
            if len(group) == 1:
                self._reduce_flat(group[0])
                if scale is not None and scale != 1.0:
                    group[0] *= scale
                continue
            total = sum(g.size for g in group)
            bucket = self._bucket(total, group[0].dtype)
            stream = C.c_void_p(current_stream())
            item = group[0].dtype.itemsize
            off = 0
            for g in group:
                check(lib.mgb_memcpy_d2d(C.c_void_p(bucket.ptr + off * item), C.c_void_p(g.ptr),
                                         C.c_size_t(g.nbytes), stream))
                off += g.size
            flat = DeviceArray((total,), group[0].dtype, _ptr=bucket.ptr, _base=bucket)
            self._reduce_flat(flat)
            if scale is not None and scale != 1.0:
                flat *= scale
            off = 0
            for g in group:
                check(lib.mgb_memcpy_d2d(C.c_void_p(g.ptr), C.c_void_p(bucket.ptr + off * item),
                                         C.c_size_t(g.nbytes), stream))
                off += g.size

    def _bucket(self, n: int, dtype) -> DeviceArray:
        key = np.dtype(dtype).str
        buf = self._buckets.get(key)
        if buf is None or buf.size < n:
            buf = DeviceArray((max(n, 1 << 16),), dtype)
            self._buckets[key] = buf
        return buf
