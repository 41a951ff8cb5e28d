"""``mygrad.nnet.layers`` entry points served by the B200 path (layers/__init__.py:1-5)."""
from .batchnorm import BatchNorm, batchnorm
from .conv import ConvND, conv_nd
from .conv_pool import ConvPoolND, conv_pool
from .pooling import MaxPoolND, MeanPoolND, max_pool, mean_pool

__all__ = ["conv_nd", "max_pool", "mean_pool", "conv_pool", "batchnorm", "ConvND", "MaxPoolND", "MeanPoolND",
           "ConvPoolND", "BatchNorm"]
