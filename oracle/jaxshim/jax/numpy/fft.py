"""jax.numpy.fft subset of the shim (test infrastructure)."""
import torch

from .._array import Array, as_tensor


def fftn(x):
    return Array(torch.fft.fftn(as_tensor(x)))
