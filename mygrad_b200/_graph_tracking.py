"""TRACK_GRAPH switch and the `no_autodiff` context manager / decorator.

Same semantics as src/mygrad/_utils/graph_tracking.py:17-91: a module-level
boolean consulted by Tensor._op and Tensor.backward; `no_autodiff` clears it
for the duration of a block (or a decorated call) and restores the old value.
MEM_GUARD write-locking (lock_management.py) is not reproduced: device buffers
have no writeable flag (documented divergence).
"""
from functools import wraps

__all__ = ["no_autodiff", "TRACK_GRAPH"]

TRACK_GRAPH = True


class _NoAutoDiff:
    """Usable as ``with no_autodiff:`` and as ``@no_autodiff``."""

    def __init__(self):
        self._stack = []

    def __enter__(self):
        global TRACK_GRAPH
        self._stack.append(TRACK_GRAPH)
        TRACK_GRAPH = False
        return self

    def __exit__(self, exc_type, exc, tb):
        global TRACK_GRAPH
        TRACK_GRAPH = self._stack.pop()
        return False

    def __call__(self, func):
        @wraps(func)
        def wrapper(*args, **kwargs):
            with self:
                return func(*args, **kwargs)

        return wrapper


no_autodiff = _NoAutoDiff()
