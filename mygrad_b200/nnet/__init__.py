"""Neural-network layers of the device path (mirrors ``mygrad.nnet``)."""
from . import layers
from .layers import conv_nd, max_pool, mean_pool

__all__ = ["layers", "conv_nd", "max_pool", "mean_pool"]
