"""jax.numpy subset of the shim (test infrastructure): exactly the functions the reference's
admp/{multipole,spatial,pairwise,pme,recip,disp_pme}.py call, with NumPy semantics, on torch float64."""
import math

import numpy as _np
import torch

from .._array import Array, F64, as_tensor, build, raw
from . import linalg, fft   # noqa: F401

pi = math.pi
newaxis = None


def _t(x):
    return as_tensor(x)


def _f(x):
    t = as_tensor(x)
    return t if (t.is_floating_point() or t.is_complex()) else t.to(F64)


def _shape(shape):
    if isinstance(shape, (tuple, list)):
        return tuple(int(s) for s in shape)
    return (int(shape),)


def array(obj, dtype=None):
    if isinstance(obj, Array):
        return Array(obj.t.clone())
    if isinstance(obj, (list, tuple)):
        return Array(build(obj))
    return Array(as_tensor(obj).clone())


def asarray(obj, dtype=None):
    if isinstance(obj, Array):
        return obj
    return array(obj)


def zeros(shape, dtype=None):
    return Array(torch.zeros(_shape(shape), dtype=F64))


def empty(shape, dtype=None):
    return Array(torch.zeros(_shape(shape), dtype=F64))


def zeros_like(x):
    return Array(torch.zeros_like(_t(x)))


def eye(n):
    return Array(torch.eye(int(n), dtype=F64))


def arange(*args):
    return Array(torch.arange(*[int(a) for a in args], dtype=torch.int64))


def linspace(start, stop, num):
    return Array(torch.linspace(float(start), float(stop), int(num), dtype=F64))


def meshgrid(*xs, indexing='xy'):
    return [Array(g) for g in torch.meshgrid(*[_t(x) for x in xs], indexing=indexing)]


def _reduce(fn, x, axis, keepdims):
    t = _t(x)
    if axis is None:
        return Array(fn(t))
    return Array(fn(t, dim=axis, keepdim=keepdims))


def sum(x, axis=None, keepdims=False):   # noqa: A001
    return _reduce(torch.sum, x, axis, keepdims)


def prod(x, axis=None, keepdims=False):
    return _reduce(torch.prod, x, axis, keepdims)


def max(x, axis=None, keepdims=False):   # noqa: A001
    t = _t(x)
    if axis is None:
        return Array(torch.max(t))
    return Array(torch.amax(t, dim=axis, keepdim=keepdims))


def _unary(fn, floating=True):
    def f(x):
        return Array(fn(_f(x) if floating else _t(x)))
    return f


exp = _unary(torch.exp)
log = _unary(torch.log)
sqrt = _unary(torch.sqrt)
cos = _unary(torch.cos)
floor = _unary(torch.floor)
ceil = _unary(torch.ceil)
round = _unary(torch.round)   # noqa: A001
abs = _unary(torch.abs, False)   # noqa: A001
real = _unary(torch.real, False)
imag = _unary(torch.imag, False)
logical_not = _unary(torch.logical_not, False)


def logical_and(a, b):
    return Array(torch.logical_and(_t(a), _t(b)))


def logical_or(a, b):
    return Array(torch.logical_or(_t(a), _t(b)))


def where(cond, a, b):
    a, b = _t(a), _t(b)
    dt = torch.promote_types(a.dtype, b.dtype)
    return Array(torch.where(_t(cond), a.to(dt), b.to(dt)))


def mod(a, b):
    return Array(torch.remainder(_t(a), _t(b)))


def stack(seq, axis=0):
    return Array(torch.stack([_t(s) for s in seq], dim=axis))


def hstack(seq):
    ts = [_t(s) for s in seq]
    dt = ts[0].dtype
    for p in ts[1:]:
        dt = torch.promote_types(dt, p.dtype)
    return Array(torch.hstack([p.to(dt) for p in ts]))


def dot(a, b):
    a, b = _t(a), _t(b)
    dt = torch.promote_types(a.dtype, b.dtype)
    return Array(torch.matmul(a.to(dt), b.to(dt)))


def einsum(spec, *ops):
    ts = [_t(o) for o in ops]
    dt = ts[0].dtype
    for p in ts[1:]:
        dt = torch.promote_types(dt, p.dtype)
    return Array(torch.einsum(spec, *[p.to(dt) for p in ts]))


def cross(a, b):
    return Array(torch.linalg.cross(_t(a), _t(b), dim=-1))


def trace(x, axis1=0, axis2=1):
    return Array(torch.diagonal(_t(x), dim1=axis1, dim2=axis2).sum(-1))


def swapaxes(x, a, b):
    return Array(_t(x).transpose(a, b))


def roll(x, shift):
    return Array(torch.roll(_t(x), int(shift)))


def piecewise(x, condlist, funclist):
    """jnp.piecewise: under jit/vmap jax lowers it to a select over all branches; the same here
    (every branch is evaluated on the whole input, gradients flow through the selected one)."""
    xt = _t(x)
    out = torch.zeros_like(xt) if len(funclist) == len(condlist) else None
    if out is None:
        d = funclist[-1]
        out = _t(d(Array(xt))) if callable(d) else _t(d)
        out = out + torch.zeros_like(xt)
    for cond, fn in zip(condlist, funclist):
        val = _t(fn(Array(xt))) if callable(fn) else _t(fn)
        out = torch.where(_t(cond), val.to(out.dtype), out)
    return Array(out)


def set_printoptions(**kw):
    pass
