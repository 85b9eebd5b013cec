"""Immutable array type of the jax API shim (TEST INFRASTRUCTURE, see ../README.md).

``Array`` wraps a ``torch.Tensor`` (float64 / int64 / bool / complex128 on the CPU) and gives it the
slice of the ``jax.Array`` surface the reference's ``admp/*.py`` uses: NumPy operator semantics,
``.T`` on any rank, ``.dot``, ``.astype``, functional ``.at[idx].set/add``, and NO in-place
operators (``x += y`` rebinds, as in JAX).  torch.autograd supplies ``jax.grad``; ``torch.vmap``
supplies ``jax.vmap`` (the wrapped tensor may be a BatchedTensor).
"""
import operator

import numpy as np
import torch

F64 = torch.float64


def raw(x):
    """Array / ndarray / python scalar / nested list -> torch tensor (python ints stay ints)."""
    if isinstance(x, Array):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t.to(F64) if t.dtype in (torch.float32, torch.float16) else t
    if isinstance(x, (bool, np.bool_)):
        return bool(x)
    if isinstance(x, (int, np.integer)):
        return int(x)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=F64)
    if isinstance(x, complex):
        return torch.tensor(x, dtype=torch.complex128)
    if isinstance(x, (list, tuple)):
        return build(x)
    raise TypeError('jax shim: cannot convert %r' % type(x))


def as_tensor(x):
    t = raw(x)
    if isinstance(t, bool):
        return torch.tensor(t)
    if isinstance(t, int):
        return torch.tensor(t, dtype=torch.int64)
    return t


def build(obj):
    """jnp.array of a (nested) sequence whose leaves may be Arrays (possibly batched)."""
    if not isinstance(obj, (list, tuple)):
        return as_tensor(obj)
    parts = [build(o) for o in obj]
    if not parts:
        return torch.zeros(0, dtype=F64)
    dt = parts[0].dtype
    for p in parts[1:]:
        dt = torch.promote_types(dt, p.dtype)
    parts = torch.broadcast_tensors(*[p.to(dt) for p in parts])
    return torch.stack(list(parts), dim=0)


def wrap(t):
    if isinstance(t, torch.Tensor):
        return Array(t)
    if isinstance(t, (tuple, list)):
        return type(t)(wrap(u) for u in t)
    return t


def _index(idx):
    if isinstance(idx, tuple):
        return tuple(_index(i) for i in idx)
    if isinstance(idx, Array):
        return idx.t
    if isinstance(idx, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(idx))
    if isinstance(idx, (np.integer,)):
        return int(idx)
    if isinstance(idx, list):
        return torch.as_tensor(idx)
    return idx


def _binary(op, reflected=False):
    def f(self, other):
        if other is None:
            return NotImplemented
        o = raw(other)
        a = self.t
        if isinstance(o, torch.Tensor) and o.dtype != a.dtype and not (o.is_complex() or a.is_complex()):
            # NumPy promotion: int (x) float -> float64, never torch's default float32
            if o.is_floating_point() != a.is_floating_point():
                if a.dtype != torch.bool and o.dtype != torch.bool:
                    a, o = a.to(F64), o.to(F64)
        if op is operator.truediv:         # NumPy: int / int -> float64 (torch would give float32)
            if not (a.is_floating_point() or a.is_complex()):
                a = a.to(F64)
            if isinstance(o, torch.Tensor) and not (o.is_floating_point() or o.is_complex()):
                o = o.to(F64)
        return Array(op(o, a) if reflected else op(a, o))
    return f


class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


class _AtRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _functional_int(self, value, add):
        # vmap-safe path: 1-D array, python int index (the bVec recursion of admp/pme.py:293-300)
        a = self.arr.t
        v = as_tensor(value)
        k0 = self.idx % a.shape[0]
        parts = []
        for k in range(a.shape[0]):
            if k == k0:
                parts.append((a[k] + v) if add else v + 0 * a[k])
            else:
                parts.append(a[k])
        dt = parts[0].dtype
        for p in parts[1:]:
            dt = torch.promote_types(dt, p.dtype)
        return Array(torch.stack([p.to(dt) for p in torch.broadcast_tensors(*parts)], dim=0))

    def set(self, value):
        if isinstance(self.idx, (int, np.integer)) and self.arr.t.dim() == 1:
            return self._functional_int(value, add=False)
        out = self.arr.t.clone()
        v = as_tensor(value)
        idx = _index(self.idx)
        if out[idx].numel() == 0:      # jax: an empty scatter is a no-op whatever the update's shape
            return Array(out)
        out[idx] = v.to(out.dtype)
        return Array(out)

    def add(self, value):
        if isinstance(self.idx, (int, np.integer)) and self.arr.t.dim() == 1:
            return self._functional_int(value, add=True)
        idx = _index(self.idx)
        out = self.arr.t.clone()
        v = as_tensor(value).to(out.dtype)
        advanced = isinstance(idx, tuple) and all(isinstance(i, torch.Tensor) for i in idx)
        if advanced:
            out.index_put_(tuple(i.long() for i in idx), v, accumulate=True)   # duplicates accumulate (recip.py:327)
        else:
            out[idx] = out[idx] + v
        return Array(out)


class Array:
    __array_priority__ = 10000
    __array_ufunc__ = None          # numpy binary operators defer to the reflected methods below
    __slots__ = ('t',)

    def __init__(self, t):
        self.t = t

    # --- introspection
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def size(self):
        return self.t.numel()

    @property
    def T(self):
        return Array(self.t.permute(*reversed(range(self.t.dim()))))

    @property
    def at(self):
        return _AtIndexer(self)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        for k in range(self.t.shape[0]):
            yield Array(self.t[k])

    def __bool__(self):
        return bool(self.t)

    def __int__(self):
        return int(self.t)

    def __float__(self):
        return float(self.t)

    def __index__(self):
        return int(self.t)

    def __repr__(self):
        return 'ShimArray(%r)' % (self.t,)

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    # --- indexing
    def __getitem__(self, idx):
        return Array(self.t[_index(idx)])

    # --- arithmetic (no in-place variants on purpose)
    __add__ = _binary(operator.add)
    __radd__ = _binary(operator.add, True)
    __sub__ = _binary(operator.sub)
    __rsub__ = _binary(operator.sub, True)
    __mul__ = _binary(operator.mul)
    __rmul__ = _binary(operator.mul, True)
    __truediv__ = _binary(operator.truediv)
    __rtruediv__ = _binary(operator.truediv, True)
    __floordiv__ = _binary(operator.floordiv)
    __mod__ = _binary(operator.mod)
    __pow__ = _binary(operator.pow)
    __rpow__ = _binary(operator.pow, True)
    __lt__ = _binary(operator.lt)
    __le__ = _binary(operator.le)
    __gt__ = _binary(operator.gt)
    __ge__ = _binary(operator.ge)
    __eq__ = _binary(operator.eq)
    __ne__ = _binary(operator.ne)
    __and__ = _binary(torch.logical_and)
    __or__ = _binary(torch.logical_or)
    __hash__ = None

    def __neg__(self):
        return Array(-self.t)

    def __pos__(self):
        return self

    def __abs__(self):
        return Array(torch.abs(self.t))

    def __invert__(self):
        return Array(torch.logical_not(self.t))

    def __matmul__(self, other):
        return self.dot(other)

    # --- methods
    def dot(self, other):
        a, b = self.t, as_tensor(other)
        dt = torch.promote_types(a.dtype, b.dtype)
        return Array(torch.matmul(a.to(dt), b.to(dt)))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self.t.reshape(tuple(int(s) for s in shape)))

    def flatten(self):
        return Array(self.t.reshape(-1))

    def astype(self, dtype):
        if dtype in (int, 'int', np.int64, torch.int64):
            return Array(self.t.to(torch.int64))
        if dtype in (float, 'float', np.float64, torch.float64):
            return Array(self.t.to(F64))
        if dtype in (bool, np.bool_, torch.bool):
            return Array(self.t.to(torch.bool))
        raise TypeError('jax shim: astype(%r)' % (dtype,))

    def swapaxes(self, a, b):
        return Array(self.t.transpose(a, b))

    def sum(self, axis=None, keepdims=False):
        if axis is None:
            return Array(self.t.sum())
        return Array(self.t.sum(dim=axis, keepdim=keepdims))
