"""jax.scipy.special.{erf, erfc} on torch.special (test infrastructure)."""
import torch

from .._array import Array, as_tensor


def erf(x):
    return Array(torch.special.erf(as_tensor(x)))


def erfc(x):
    return Array(torch.special.erfc(as_tensor(x)))
