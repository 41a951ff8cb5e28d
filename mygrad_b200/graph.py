"""StepGraph: record one training step as a CUDA graph and replay it with a single launch.

The reference runs NumPy eagerly and has no counterpart; this is for launch-bound steps (BASELINE config 1:
eight kernels of a few microseconds behind ~150 us of Python and driver work per step).  The step runs on a
private stream: twice eagerly, so that the allocator's exact-size cache holds every block the step needs, then
once more under ``cudaStreamBeginCapture``.  Replays write into the same buffers: the objects the captured call
returned (``graph.result``) are refreshed by every ``replay()``; inputs are whatever device arrays the step
closes over — update them in place (``DeviceArray.set_from_host``) between replays.

Not capturable: anything that synchronises (``.get()``, ``.item()``), steps whose buffer sizes change from call to
call, and the peer-memory collectives of ``DataParallel`` (their epoch counters are kernel arguments).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable

from . import device_array as _da
from ._lib import check, lib

__all__ = ["StepGraph"]


class StepGraph:
    def __init__(self, step: Callable[[], Any], warmup: int = 2):
        s = C.c_void_p()
        check(lib.mgb_stream_create(C.byref(s)))
        self.stream = int(s.value)
        self._exec = C.c_void_p()
        self.kernel_nodes = 0
        prev = _da.current_stream()
        check(lib.mgb_stream_synchronize(C.c_void_p(prev)))      # inputs written on the caller's stream are ready
        _da._GRAPH_STREAMS.add(self.stream)
        _da.set_stream(self.stream)
        try:
            keep = None
            for _ in range(max(1, warmup)):
                keep = step()                                     # fills the block cache of this stream
            check(lib.mgb_stream_synchronize(s))
            del keep
            check(lib.mgb_graph_begin(s))
            _da._capturing = True
            try:
                self.result = step()
            finally:
                _da._capturing = False
                n = C.c_int(0)
                rc = lib.mgb_graph_end(s, C.byref(self._exec), C.byref(n))
            check(rc)
            self.kernel_nodes = int(n.value)
        except Exception:
            _da._GRAPH_STREAMS.discard(self.stream)
            raise
        finally:
            _da.set_stream(prev)

    def replay(self) -> Any:
        """enqueues the recorded step on the graph's stream; returns the (refreshed) result objects"""
        check(lib.mgb_graph_launch(self._exec, C.c_void_p(self.stream)))
        return self.result

    def synchronize(self) -> None:
        check(lib.mgb_stream_synchronize(C.c_void_p(self.stream)))

    def close(self) -> None:
        if self._exec:
            lib.mgb_stream_synchronize(C.c_void_p(self.stream))
            lib.mgb_graph_destroy(self._exec)
            self._exec = C.c_void_p()
            self.result = None
            _da._GRAPH_STREAMS.discard(self.stream)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
