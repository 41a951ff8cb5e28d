"""``mygrad.nnet.losses`` entries served on the device (only what config 5 needs)."""
from .softmax_crossentropy import SoftmaxCrossEntropy, softmax_crossentropy

__all__ = ["softmax_crossentropy", "SoftmaxCrossEntropy"]
