"""Reciprocal-space calculator generator - drop-in surface of admp/recip.py:21-462.

``generate_pme_recip(Ck_fn, kappa, gamma, pme_order, K1, K2, K3, lmax)`` returns
``pme_recip(positions, box, Q)``.  ``Ck_fn`` is one of the markers ``Ck_1``, ``Ck_6``,
``Ck_8``, ``Ck_10`` naming the influence function evaluated inside the convolution kernel.
Differentiable: positions, box, Q.
"""
import torch

from . import _lib
from ._ctx import Context, to_dev

sqrt_pi = 1.7724538509055159     # admp/recip.py:19


class InfluenceFunction:
    def __init__(self, name, kind, gamma):
        self.name, self.kind, self.gamma = name, kind, gamma


Ck_1 = InfluenceFunction('Ck_1', _lib.CK_COULOMB, False)     # admp/recip.py:434
Ck_6 = InfluenceFunction('Ck_6', _lib.CK_DISP6, True)        # admp/recip.py:437
Ck_8 = InfluenceFunction('Ck_8', _lib.CK_DISP8, True)        # admp/recip.py:445
Ck_10 = InfluenceFunction('Ck_10', _lib.CK_DISP10, True)     # admp/recip.py:454


class _RecipFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cx, kind, lmax, Kvec, positions, box, Q):
        n = ctx.needs_input_grad
        ctx.kvec_ref = getattr(cx, 'kvec_ref', False)
        flags = 0
        if n[4] or n[5] or n[6]:
            flags |= _lib.WANT_GRAD
        if n[5]:
            flags |= _lib.WANT_VIRIAL
        ctx.Kvec = Kvec
        na, dt, dev = positions.shape[0], cx.dtype, cx.device
        nh = (lmax + 1) ** 2
        p = _lib.ptr
        scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
        dpos = torch.zeros((na, 3), dtype=dt, device=dev)
        if lmax == 0:
            M, cols, stride = Q.contiguous(), 1, 1
            G = torch.zeros((na, 1), dtype=dt, device=dev)
            gs = 1
        else:
            # harmonic -> Cartesian site layout with an identity frame (lmax taken from the ctx)
            M = torch.empty((na, 10), dtype=dt, device=dev)
            _lib.check(cx.lib.admp_frames_fwd(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(Q), p(M), None, None))
            cols, stride, gs = 10, 10, 10
            G = torch.zeros((na, 10), dtype=dt, device=dev)
        _lib.check(cx.lib.admp_pme_recip(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(M), cols, stride, None, kind, 0,
                                         flags, p(dpos), p(G), gs, None, p(scal)))
        dQ = None
        if flags & _lib.WANT_GRAD:
            if lmax == 0:
                dQ = G
            else:
                dQ = torch.empty((na, nh), dtype=dt, device=dev)
                dummy = torch.zeros_like(dpos)
                _lib.check(cx.lib.admp_frames_bwd(cx.handle, _lib.stream_ptr(), p(positions), p(box), p(Q), p(G), p(dQ),
                                                  p(dummy), p(scal)))
        ctx.saved = (scal, dpos, dQ, box)
        ctx.dtype = dt
        return scal[_lib.S_E_RECIP].to(dt)

    @staticmethod
    def backward(ctx, g):
        scal, dpos, dQ, box = ctx.saved
        n, dt = ctx.needs_input_grad, ctx.dtype
        dbox = None
        if n[5]:
            # assemble dE/dbox from the accumulators exactly as virial_finalize does on the device
            inv = torch.linalg.inv(box.double())
            W = scal[_lib.S_DNSTAR:_lib.S_DNSTAR + 9].reshape(3, 3)
            tk = scal[_lib.S_TK:_lib.S_TK + 6]
            T = torch.stack([tk[0], tk[1], tk[2], tk[1], tk[3], tk[4], tk[2], tk[4], tk[5]]).reshape(3, 3)
            if ctx.kvec_ref:          # settings.KVEC_ORDER = 'reference': the reference's k table swaps mesh axes 0 and 1 (recip.py:339-341)
                T = T[[1, 0, 2]][:, [1, 0, 2]]
            nstar = (ctx.Kvec[None, :] * inv).T
            dbox = -(inv.T @ (W.T @ nstar)) - 2 * T @ inv.T - scal[_lib.S_E_RECIP] * inv.T
            dbox = (g * dbox).to(dt)
        return (None, None, None, None, g * dpos if n[4] else None, dbox, g * dQ if n[6] else None)


def generate_pme_recip(Ck_fn, kappa, gamma, pme_order, K1, K2, K3, lmax):
    """admp/recip.py:21-31.  ``gamma`` must match the influence function (Coulomb drops the
    gamma point, dispersion keeps it), as in every call site of the reference."""
    if not isinstance(Ck_fn, InfluenceFunction):
        raise TypeError('Ck_fn must be one of admp_b200.recip.Ck_1 / Ck_6 / Ck_8 / Ck_10')
    if bool(gamma) != Ck_fn.gamma:
        raise NotImplementedError('gamma=%r with %s is not a combination the reference uses' % (gamma, Ck_fn.name))
    if pme_order != 6:
        raise NotImplementedError('only pme_order = 6 is implemented (as in admp/recip.py:25)')
    if lmax > 2:
        raise NotImplementedError('l > 2 (beyond quadrupole) not supported')
    cx = Context()
    cx.set_pme(kappa, K1, K2, K3, lmax)
    from . import settings
    cx.kvec_ref = settings.KVEC_ORDER == 'reference'
    state = {'n': 0}
    Kvec = torch.tensor([float(K1), float(K2), float(K3)], dtype=torch.float64, device=cx.device)

    def pme_recip(positions, box, Q):
        positions, box, Q = (to_dev(x, cx.dtype, cx.device) for x in (positions, box, Q))
        na = positions.shape[0]
        if Q.dim() == 1:
            Q = Q[:, None]
        if Q.shape != (na, (lmax + 1) ** 2):
            raise ValueError('Q must be (%d, %d)' % (na, (lmax + 1) ** 2))
        if state['n'] != na:
            cx.set_topology(na, None, None, None)      # no frames: Q is already global
            state['n'] = na
        return _RecipFunction.apply(cx, Ck_fn.kind, lmax, Kvec, positions, box, Q)

    pme_recip._ctx = cx
    return pme_recip
