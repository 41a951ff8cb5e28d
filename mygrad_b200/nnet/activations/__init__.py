"""``mygrad.nnet.activations`` entries served on the device (only what config 5 needs)."""
from .relu import ReLu, relu

__all__ = ["relu", "ReLu"]
