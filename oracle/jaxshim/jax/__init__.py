"""Minimal `jax` API shim on torch (TEST INFRASTRUCTURE - see oracle/jaxshim/README.md).

Exists so that the UNMODIFIED reference sources /root/reference/admp/{settings,multipole,spatial,
pairwise,pme,recip,disp_pme}.py can be imported and executed in an image without JAX:
``jit`` -> identity, ``vmap`` -> torch.vmap, ``grad`` / ``value_and_grad`` -> torch.autograd,
``lax.stop_gradient`` -> detach.  float64 everywhere (the reference's PRECISION = 'double').
Only the API surface those seven files touch is provided; anything else raises AttributeError.
"""
import functools

import torch

from ._array import Array, as_tensor, raw, wrap
from . import numpy, scipy, lax, config   # noqa: F401  (submodules importable as attributes)

__shim__ = True


def jit(fun=None, static_argnums=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _is_arraylike(x):
    import numpy as np
    return isinstance(x, (Array, torch.Tensor, np.ndarray))


def vmap(fun, in_axes=0, out_axes=0):
    """jax.vmap for positional arguments. Non-array arguments (None, python scalars, bools) are
    closed over, which is what jax does for empty pytrees / static values with in_axes None."""
    def unwrap_out(o):
        if isinstance(o, Array):
            return o.t
        if isinstance(o, (tuple, list)):
            return tuple(unwrap_out(u) for u in o)
        return o

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise TypeError('jax shim vmap: in_axes %r does not match %d arguments' % (in_axes, len(args)))
        dyn_pos, dyn_val, dyn_axes = [], [], []
        for k, (a, ax) in enumerate(zip(args, axes)):
            if _is_arraylike(a):
                dyn_pos.append(k)
                dyn_val.append(as_tensor(a))
                dyn_axes.append(ax)

        def inner(*tensors):
            full = list(args)
            for k, t in zip(dyn_pos, tensors):
                full[k] = Array(t)
            return unwrap_out(fun(*full))

        if all(ax is None for ax in dyn_axes):
            return wrap(inner(*dyn_val))
        out = torch.vmap(inner, in_dims=tuple(dyn_axes), out_dims=out_axes)(*dyn_val)
        return wrap(out)
    return mapped


def _grad_impl(fun, argnums, with_value):
    single = not isinstance(argnums, (tuple, list))
    nums = (argnums,) if single else tuple(argnums)

    @functools.wraps(fun)
    def g(*args, **kwargs):
        args = list(args)
        leaves = []
        for k in nums:
            t = as_tensor(args[k]).detach().clone().requires_grad_(True)
            leaves.append(t)
            args[k] = Array(t)
        with torch.enable_grad():
            out = fun(*args, **kwargs)
            val = raw(out)
            grads = torch.autograd.grad(val, leaves, allow_unused=True)
        grads = [Array(torch.zeros_like(l) if gr is None else gr) for l, gr in zip(leaves, grads)]
        res = grads[0] if single else tuple(grads)
        if with_value:
            return Array(val.detach()), res
        return res
    return g


def grad(fun, argnums=0):
    return _grad_impl(fun, argnums, False)


def value_and_grad(fun, argnums=0):
    return _grad_impl(fun, argnums, True)
