"""Neural-network layers of the device path (mirrors ``mygrad.nnet``)."""
from . import activations, layers, losses
from .activations import relu
from .losses import softmax_crossentropy
from .layers import batchnorm, conv_nd, conv_pool, max_pool, mean_pool

__all__ = ["layers", "activations", "losses", "conv_nd", "max_pool", "mean_pool", "conv_pool", "batchnorm", "relu",
           "softmax_crossentropy"]
