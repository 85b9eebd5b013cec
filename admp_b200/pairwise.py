"""Generic pairwise-interaction driver - drop-in surface of admp/pairwise.py:45-113.

``generate_pairwise_interaction(kernel, covalent_map, static_args)`` returns
``pair_int(positions, box, pairs, mScales, *atomic_params)``.  Pair kernels are CUDA
kernels selected by the marker object passed as ``kernel``; the reference ships one,
``TT_damping_qq_c6_kernel`` (Tang-Toennies damped exchange / charge penetration / C6),
which is the one implemented.  Differentiable: positions, box, mScales, a, b, q, c.
"""
import torch

from . import _lib
from ._ctx import Context, to_dev, pairs_to_dev


# admp/pairwise.py:21-42: per-pair gathers of per-atom arrays (vmapped indexing upstream). The device kernels gather
# inside the pair loop, so these are only kept for callers that use them directly.
def distribute_scalar(params, index):
    return params[index]


def distribute_v3(pos, index):
    return pos[index]


def distribute_multipoles(multipoles, index):
    return multipoles[index]


def distribute_dispcoeff(c_list, index):
    return c_list[index]


class PairKernel:
    """Marker naming a device pair kernel and its per-atom parameter list."""

    def __init__(self, name, n_params):
        self.name, self.n_params = name, n_params

    def __call__(self, *args, **kwargs):
        raise RuntimeError('%s is evaluated on the GPU through generate_pairwise_interaction; it has no host body'
                           % self.name)


#: admp/pairwise.py:94-113 : f(dr, m, ai, aj, bi, bj, qi, qj, ci, cj)
TT_damping_qq_c6_kernel = PairKernel('TT_damping_qq_c6_kernel', 4)


class _TTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, holder, pairs, positions, box, mScales, a, b, q, c):
        n = ctx.needs_input_grad
        flags = 0
        if any(n[2:]):
            flags |= _lib.WANT_GRAD
        if n[3]:
            flags |= _lib.WANT_VIRIAL
        if any(n[4:]):
            flags |= _lib.WANT_PGRAD
        cx = holder.ctx
        na, dt, dev = cx.n_atoms, cx.dtype, cx.device
        scal = torch.empty(_lib.S_COUNT, dtype=torch.float64, device=dev)
        dpos = torch.empty((na, 3), dtype=dt, device=dev) if flags & _lib.WANT_GRAD else None
        dpar = torch.empty((4, na), dtype=dt, device=dev) if flags & _lib.WANT_PGRAD else None
        p = _lib.ptr
        _lib.check(cx.lib.admp_tt_pair(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(pairs), int(pairs.shape[0]),
                                       p(mScales), p(a), p(b), p(q), p(c), flags, p(scal), p(dpos), p(dpar)))
        ctx.saved = (scal, dpos, dpar)
        ctx.dtype = dt
        return scal[_lib.S_E_REAL].to(dt)

    @staticmethod
    def backward(ctx, g):
        scal, dpos, dpar = ctx.saved
        n, dt = ctx.needs_input_grad, ctx.dtype
        out = [None, None, g * dpos if n[2] else None,
               (g * scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3)).to(dt) if n[3] else None,
               (g * scal[_lib.S_DMSCALE:_lib.S_DMSCALE + 5]).to(dt) if n[4] else None]
        for k in range(4):
            out.append(g * dpar[k] if n[5 + k] else None)
        return tuple(out)


class _Holder:
    def __init__(self, covalent_map):
        self.ctx = Context()
        self.ctx.set_topology(int(covalent_map.shape[0]), None, None, covalent_map)


def generate_pairwise_interaction(pair_int_kernel, covalent_map, static_args=None):
    '''
    admp/pairwise.py:45-91: calculator generator for pairwise interactions.
    Output: pair_int(positions, box, pairs, mScales, *atomic_params) -> energy
    '''
    if pair_int_kernel is not TT_damping_qq_c6_kernel:
        raise NotImplementedError('only TT_damping_qq_c6_kernel has a device implementation')
    holder = _Holder(covalent_map)

    def pair_int(positions, box, pairs, mScales, *atomic_params):
        if len(atomic_params) != pair_int_kernel.n_params:
            raise TypeError('%s takes %d per-atom parameter arrays' % (pair_int_kernel.name, pair_int_kernel.n_params))
        cx = holder.ctx
        prep = lambda x: to_dev(x, cx.dtype, cx.device)
        positions, box, mScales = prep(positions), prep(box), prep(mScales)
        params = [prep(x) for x in atomic_params]
        return _TTFunction.apply(holder, pairs_to_dev(pairs, cx.device), positions, box, mScales, *params)

    pair_int._holder = holder
    return pair_int
