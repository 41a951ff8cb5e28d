"""Exception types of the graph runtime (names follow src/mygrad/errors.py:1-22)."""

__all__ = ["InvalidGradient", "InvalidBackprop", "DisconnectedView", "SkipGradient"]


class InvalidGradient(Exception):
    """An invalid gradient-value was passed to, or propagated through, a tensor"""


class InvalidBackprop(Exception):
    """Raised when trying to backpropagate through a graph that was cleared"""


class DisconnectedView(Exception):
    """Kept for name parity; views other than `reshape` are not part of the device path"""


class SkipGradient(Exception):
    """`backward_var` may raise this to skip an input (src/mygrad/_utils/__init__.py:126)"""
