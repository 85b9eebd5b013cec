"""jax.numpy.linalg subset of the shim (test infrastructure)."""
import torch

from .._array import Array, as_tensor


def inv(x):
    return Array(torch.linalg.inv(as_tensor(x)))


def det(x):
    return Array(torch.linalg.det(as_tensor(x)))


def norm(x, ord=None, axis=None, keepdims=False):
    t = as_tensor(x)
    if ord not in (None, 2):
        raise NotImplementedError('jax shim: norm ord=%r' % (ord,))
    if axis is None:
        return Array(torch.sqrt(torch.sum(t * t)))
    return Array(torch.sqrt(torch.sum(t * t, dim=axis, keepdim=keepdims)))
