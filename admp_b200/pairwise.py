"""Generic pairwise-interaction driver - drop-in surface of admp/pairwise.py:45-113.

``generate_pairwise_interaction(kernel, covalent_map, static_args)`` returns
``pair_int(positions, box, pairs, mScales, *atomic_params)`` for ANY pair kernel, as upstream:
  * kernels with a fused CUDA body (``PAIR_KERNELS``): the reference's ``TT_damping_qq_c6_kernel`` (Tang-Toennies damped
    exchange / charge penetration / C6) and ``TT_damping_qq_c6_c8_c10_kernel`` - one pass, energy + all adjoints;
  * any other callable ``kernel(dr, m, p1i, p1j, ...)`` written with element-wise tensor operations: CUDA kernels
    produce (dr, scale index) per row and push dE/d(dr) back to positions and box, the kernel body runs in between.
Differentiable: positions, box, mScales and every per-atom parameter.
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
    """A pair kernel with a fused CUDA body (one pass: energy + every adjoint). Calling it with tensors evaluates the
    same formula element-wise (what the reference's vmapped kernel does), so user code that composes or inspects kernels
    keeps working."""

    def __init__(self, name, n_params, entry, host_body):
        self.name, self.n_params, self.entry, self._host_body = name, n_params, entry, host_body

    def __call__(self, dr, m, *pair_params):
        return self._host_body(dr, m, *pair_params)


def _tt_poly(x, n):
    poly, term = torch.ones_like(x), torch.ones_like(x)
    for k in range(1, n + 1):
        term = term * x / k
        poly = poly + term
    return poly


def _tt_c6_body(dr, m, ai, aj, bi, bj, qi, qj, ci, cj):
    """admp/pairwise.py:94-113, element-wise on tensors."""
    a, b = torch.sqrt(ai * aj), torch.sqrt(bi * bj)
    br = b * dr * 1.889726878
    e = torch.exp(-br)
    return (2625.5 * a * e - 2625.5 * e * (1 + br) * (qi * qj) / br + e * _tt_poly(br, 6) * (ci * cj) / dr ** 6) * m


def _tt_c10_body(dr, m, ai, aj, bi, bj, qi, qj, c6i, c6j, c8i, c8j, c10i, c10j):
    a, b = torch.sqrt(ai * aj), torch.sqrt(bi * bj)
    br = b * dr * 1.889726878
    e = torch.exp(-br)
    disp = _tt_poly(br, 6) * (c6i * c6j) / dr ** 6 + _tt_poly(br, 8) * (c8i * c8j) / dr ** 8 + _tt_poly(br, 10) * (c10i * c10j) / dr ** 10
    return (2625.5 * a * e - 2625.5 * e * (1 + br) * (qi * qj) / br + e * disp) * m


#: admp/pairwise.py:94-113 : f(dr, m, ai, aj, bi, bj, qi, qj, ci, cj)
TT_damping_qq_c6_kernel = PairKernel('TT_damping_qq_c6_kernel', 4, 'admp_tt_pair', _tt_c6_body)
#: the same with Tang-Toennies damped C8 and C10 terms: f(dr, m, a.., b.., q.., c6.., c8.., c10..)  
TT_damping_qq_c6_c8_c10_kernel = PairKernel('TT_damping_qq_c6_c8_c10_kernel', 6, 'admp_tt_pair_c10', _tt_c10_body)

#: device kernels by name
PAIR_KERNELS = {k.name: k for k in (TT_damping_qq_c6_kernel, TT_damping_qq_c6_c8_c10_kernel)}


class _FusedPairFunction(torch.autograd.Function):
    """pair_int for a kernel with a fused CUDA body: E, dE/dpositions, dE/dbox, dE/dmScales and dE/d(per-atom parameters)."""

    @staticmethod
    def forward(ctx, holder, kernel, pairs, positions, box, mScales, *params):
        n = ctx.needs_input_grad
        flags = 0
        if any(n[3:]):
            flags |= _lib.WANT_GRAD
        if n[4]:
            flags |= _lib.WANT_VIRIAL
        if any(n[5:]):
            flags |= _lib.WANT_PGRAD
        cx = holder.ctx
        na, dt, dev = cx.n_atoms, cx.dtype, cx.device
        scal = torch.empty(_lib.S_COUNT, dtype=torch.float64, device=dev)
        dpos = torch.empty((na, 3), dtype=dt, device=dev) if flags & _lib.WANT_GRAD else None
        dpar = torch.empty((kernel.n_params, na), dtype=dt, device=dev) if flags & _lib.WANT_PGRAD else None
        p = _lib.ptr
        fn = getattr(cx.lib, kernel.entry)
        _lib.check(fn(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(pairs), int(pairs.shape[0]), p(mScales),
                      *[p(x) for x in params], flags, p(scal), p(dpos), p(dpar)))
        ctx.saved = (scal, dpos, dpar)
        ctx.dtype = dt
        ctx.n_params = kernel.n_params
        return scal[_lib.S_E_REAL].to(dt)

    @staticmethod
    def backward(ctx, g):
        scal, dpos, dpar = ctx.saved
        n, dt = ctx.needs_input_grad, ctx.dtype
        out = [None, None, None, g * dpos if n[3] else None,
               (g * scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3)).to(dt) if n[4] else None,
               (g * scal[_lib.S_DMSCALE:_lib.S_DMSCALE + 5]).to(dt) if n[5] else None]
        for k in range(ctx.n_params):
            out.append(g * dpar[k] if n[6 + k] else None)
        return tuple(out)


class _PairGeometry(torch.autograd.Function):
    """rows -> (dr, scale index): minimum-image distances by a CUDA kernel, and the CUDA adjoint that turns dE/d(dr)
    into dE/dpositions and the image-shift part of dE/dbox (admp_pair_geometry / admp_pair_geometry_bwd)."""

    @staticmethod
    def forward(ctx, holder, pairs, positions, box):
        cx = holder.ctx
        rows = int(pairs.shape[0])
        dr = torch.empty(rows, dtype=cx.dtype, device=cx.device)
        sidx = torch.empty(rows, dtype=torch.int32, device=cx.device)
        p = _lib.ptr
        _lib.check(cx.lib.admp_pair_geometry(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(pairs), rows, p(dr), p(sidx)))
        ctx.holder, ctx.pairs = holder, pairs
        ctx.save_for_backward(positions, box)
        ctx.mark_non_differentiable(sidx)
        return dr, sidx

    @staticmethod
    def backward(ctx, g_dr, _g_sidx):
        positions, box = ctx.saved_tensors
        cx = ctx.holder.ctx
        n = ctx.needs_input_grad
        dpos = torch.empty_like(positions)
        scal = torch.empty(_lib.S_COUNT, dtype=torch.float64, device=cx.device)
        p = _lib.ptr
        g = g_dr.contiguous().to(cx.dtype)
        _lib.check(cx.lib.admp_pair_geometry_bwd(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(ctx.pairs), int(ctx.pairs.shape[0]),
                                                 p(g), _lib.WANT_VIRIAL if n[3] else 0, p(dpos), p(scal)))
        dbox = scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3).to(positions.dtype) if n[3] else None
        return None, None, dpos if n[2] else None, dbox


class _Holder:
    def __init__(self, covalent_map):
        self.ctx = Context()
        self.ctx.set_topology(int(covalent_map.shape[0]), None, None, covalent_map)


def generate_pairwise_interaction(pair_int_kernel, covalent_map, static_args=None):
    '''
    admp/pairwise.py:45-91: calculator generator for pairwise interactions.
    Output: pair_int(positions, box, pairs, mScales, *atomic_params) -> energy

    pair_int_kernel is either one of the kernels with a fused CUDA body (PAIR_KERNELS: the reference's
    TT_damping_qq_c6_kernel, and TT_damping_qq_c6_c8_c10_kernel) or ANY callable
        kernel(dr, m, p1i, p1j, p2i, p2j, ...) -> per-pair energy
    written with element-wise tensor operations (the reference's kernels are the same functions written for one pair and
    vmapped). Generic kernels run between two CUDA kernels: rows -> (dr, scale index) and dE/d(dr) -> dE/dpositions, dE/dbox;
    the per-atom parameters are gathered per pair end exactly as pairwise.py:79-84 does. Rows with pairs[:,0] >= pairs[:,1]
    (padding) contribute nothing. Everything passed positionally is differentiable.
    '''
    holder = _Holder(covalent_map)
    fused = isinstance(pair_int_kernel, PairKernel)
    if not fused and not callable(pair_int_kernel):
        raise TypeError('pair_int_kernel must be a PairKernel or a callable (dr, m, *pair_params) -> energy')

    def pair_int(positions, box, pairs, mScales, *atomic_params):
        cx = holder.ctx
        prep = lambda x: to_dev(x, cx.dtype, cx.device)
        positions, box, mScales = prep(positions), prep(box), prep(mScales)
        params = [prep(x) for x in atomic_params]
        na = cx.n_atoms
        if tuple(positions.shape) != (na, 3) or tuple(box.shape) != (3, 3) or mScales.shape != (5,):
            raise ValueError('positions must be (%d, 3), box (3, 3) and mScales must hold exactly 5 entries' % na)
        for x in params:
            if x.shape != (na,):
                raise ValueError('per-atom parameters must be (%d,)' % na)
        pr = pairs_to_dev(pairs, cx.device)
        if fused:
            if len(params) != pair_int_kernel.n_params:
                raise TypeError('%s takes %d per-atom parameter arrays' % (pair_int_kernel.name, pair_int_kernel.n_params))
            return _FusedPairFunction.apply(holder, pair_int_kernel, pr, positions, box, mScales, *params)
        if int(pr.shape[0]) == 0:
            return positions.sum() * 0
        dr, sidx = _PairGeometry.apply(holder, pr, positions, box)
        live = sidx >= 0
        rows = torch.nonzero(live, as_tuple=False).squeeze(1)            # pairwise.py:60: only i < j rows are evaluated
        i, j = pr[rows, 0].long(), pr[rows, 1].long()
        m = mScales[sidx[rows].long()]
        pair_params = []
        for x in params:                                                 # pairwise.py:79-84
            pair_params.append(x[i])
            pair_params.append(x[j])
        return torch.sum(pair_int_kernel(dr[rows], m, *pair_params))

    pair_int._holder = holder
    return pair_int
