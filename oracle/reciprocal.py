"""Oracle (test infrastructure): order-6 B-spline multipolar PME reciprocal energy.

Restates admp/recip.py:21-462.  Deliberate, documented differences from the
reference:
  * A5 - k-vectors use the natural ordering (k[idx0], k[idx1], k[idx2]) of the
    flattened ``fftn`` output; the reference's ``meshgrid(kz, kx, ky)`` permutes them,
    which is harmless only for K1=K2=K3 on a cubic box (every shipped example, where
    both orderings give identical energies) and inconsistent otherwise.
  * A6 - the B-spline derivative Jacobian uses the chain-rule-correct index order;
    ``jacobian='reference'`` reproduces the reference's transposed form.  Both are
    identical for orthorhombic boxes (and for their diagonal dE/dbox entries).
  * A23 - M6, M6', M6'' are evaluated from the truncated-power definition instead of
    the hard-coded piecewise polynomials (recip.py:80-137); the two agree to 4e-13.
"""
import math

import numpy as np
import torch

from . import DIELECTRIC

SQRT_PI = 1.7724538509055159       # recip.py:19
ORDER = 6

_binom6 = [1.0, -6.0, 15.0, -20.0, 15.0, -6.0, 1.0]


def _tp(u, power, fact):
    """sum_k (-1)^k C(6,k) (u-k)_+^power / fact, for u in [0, 6)."""
    out = torch.zeros_like(u)
    for k in range(6):
        out = out + _binom6[k] * torch.clamp(u - k, min=0.0) ** power
    return out / fact


def bspline(u):          # recip.py:80-98
    return _tp(u, 5, 120.0)


def bspline_prime(u):    # recip.py:101-118
    return _tp(u, 4, 24.0)


def bspline_prime2(u):   # recip.py:121-137
    return _tp(u, 3, 6.0)


def _stencil_shifts():
    """recip.py:27-29: the 216 integer offsets {-3..2}^3 (order is immaterial: the
    same table indexes both the weights and the mesh points)."""
    r = torch.arange(-ORDER // 2, ORDER // 2)
    g = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), dim=-1).reshape(-1, 3)
    return g


def spread(positions, box, Q, K, lmax, jacobian='correct'):
    """recip.py:368-392 (spread_Q) with its helpers :37-329.  Returns the real mesh."""
    N = torch.tensor([float(K[0]), float(K[1]), float(K[2])], dtype=positions.dtype)
    Nstar = (N.reshape(1, 3) * torch.linalg.inv(box)).T              # :55
    Rm = positions @ Nstar.T                                          # :75
    m0 = torch.ceil(Rm.detach())
    u0 = (m0 - Rm) + ORDER / 2                                        # :77
    sh = _stencil_shifts()
    na = positions.shape[0]
    u = u0[:, None, :] + sh[None].to(positions.dtype)                 # (na,216,3)  :236
    M = bspline(u)
    theta = M[..., 0] * M[..., 1] * M[..., 2]                         # :152
    n_harm = (lmax + 1) ** 2
    cols = [theta]
    if lmax >= 1:
        Mp = bspline_prime(u)
        div = torch.stack([Mp[..., 0] * M[..., 1] * M[..., 2],
                           Mp[..., 1] * M[..., 2] * M[..., 0],
                           Mp[..., 2] * M[..., 0] * M[..., 1]], dim=-1)        # :169-173
        J = -Nstar
        if jacobian == 'reference':
            tp = torch.einsum('ij,akj->aki', J, div)                            # :177 as written
        else:
            tp = torch.einsum('ji,akj->aki', J, div)                            # chain rule du_j/dx_i
        cols += [tp[..., 2], tp[..., 0], tp[..., 1]]                            # :246-252
    if lmax >= 2:
        Mpp = bspline_prime2(u)
        d = [[None] * 3 for _ in range(3)]
        d[0][0] = Mpp[..., 0] * M[..., 1] * M[..., 2]
        d[1][1] = Mpp[..., 1] * M[..., 0] * M[..., 2]
        d[2][2] = Mpp[..., 2] * M[..., 0] * M[..., 1]
        d[0][1] = d[1][0] = Mp[..., 0] * Mp[..., 1] * M[..., 2]
        d[0][2] = d[2][0] = Mp[..., 0] * Mp[..., 2] * M[..., 1]
        d[1][2] = d[2][1] = Mp[..., 1] * Mp[..., 2] * M[..., 0]
        D = torch.stack([torch.stack(row, dim=-1) for row in d], dim=-2)        # (na,216,3,3)
        if jacobian == 'reference':
            t2 = torch.einsum('im,jn,akmn->akij', J, J, D)                      # :212 as written
        else:
            t2 = torch.einsum('mi,nj,akmn->akij', J, J, D)
        rt3 = math.sqrt(3.0)                                                    # :263
        tr = t2[..., 0, 0] + t2[..., 1, 1] + t2[..., 2, 2]
        cols += [(3 * t2[..., 2, 2] - tr) / 2, rt3 * t2[..., 0, 2], rt3 * t2[..., 1, 2],
                 rt3 / 2 * (t2[..., 0, 0] - t2[..., 1, 1]), rt3 * t2[..., 0, 1]]   # :266-270
    harm = torch.stack(cols, dim=-1)                                            # (na,216,n_harm)
    Qw = Q[:, :n_harm].clone()
    if lmax >= 2:
        Qw = torch.cat([Q[:, :4], Q[:, 4:9] / 3], dim=1)                        # :300-305
    per_atom = torch.sum(Qw[:, None, :] * harm, dim=2)                          # :307
    Ki = torch.tensor([int(K[0]), int(K[1]), int(K[2])])
    idx = torch.remainder(m0.long()[:, None, :] + sh[None], Ki)                 # :324
    mesh = torch.zeros(int(K[0]), int(K[1]), int(K[2]), dtype=positions.dtype)
    mesh = mesh.index_put((idx[..., 0], idx[..., 1], idx[..., 2]), per_atom, accumulate=True)   # :328
    return mesh


def kpts_int_1d(n):
    """recip.py:339: roll(arange(-(n-1)//2, (n+1)//2), -(n-1)//2) with Python's
    precedence ``(-(n-1))//2`` -> [0, 1, ..., -2, -1]; even n: Nyquist is negative."""
    lo = (-(n - 1)) // 2
    k = np.arange(lo, (n + 1) // 2)
    return np.roll(k, lo)


def theta_k_1d(n, dtype=torch.float64):
    """recip.py:400-408: sum_{m=-2..2} M6(m+3) cos(2 pi m k / n) per dimension."""
    m = torch.arange(-ORDER // 2 + 1, ORDER // 2, dtype=dtype)       # -2..2
    w = bspline(m + ORDER / 2)
    k = torch.as_tensor(kpts_int_1d(n), dtype=dtype)
    return torch.sum(w[:, None] * torch.cos(2 * math.pi * m[:, None] * k[None, :] / n), dim=0)


def Ck_1(ksq, kappa, V):     # recip.py:434-435
    return 2 * math.pi / V / ksq * torch.exp(-ksq / 4 / kappa**2)


def Ck_6(ksq, kappa, V):     # recip.py:437-443
    x2 = ksq / 4 / kappa**2
    x = torch.sqrt(x2)
    f = (1 - 2 * x2) * torch.exp(-x2) + 2 * x2 * x * SQRT_PI * torch.erfc(x)
    return SQRT_PI * math.pi / 2 / V * kappa**3 * f / 3


def Ck_8(ksq, kappa, V):     # recip.py:445-452
    x2 = ksq / 4 / kappa**2
    x = torch.sqrt(x2)
    x4 = x2 * x2
    f = (3 - 2 * x2 + 4 * x4) * torch.exp(-x2) - 4 * x4 * x * SQRT_PI * torch.erfc(x)
    return SQRT_PI * math.pi / 2 / V * kappa**5 * f / 45


def Ck_10(ksq, kappa, V):    # recip.py:454-462
    x2 = ksq / 4 / kappa**2
    x = torch.sqrt(x2)
    x4 = x2 * x2
    x6 = x4 * x2
    f = (15 - 6 * x2 + 4 * x4 - 8 * x6) * torch.exp(-x2) + 8 * x6 * x * SQRT_PI * torch.erfc(x)
    return SQRT_PI * math.pi / 2 / V * kappa**7 * f / 1260


CK = {1: Ck_1, 6: Ck_6, 8: Ck_8, 10: Ck_10}


def _safe_sqrt_ksq(ksq):
    return ksq


# Module-level defaults of the two documented divergences; tests flip them to 'reference' to reproduce the
# reference's dE/dbox entry by entry (tests/test_reference_source.py).
DEFAULTS = {'jacobian': 'correct', 'korder': 'natural'}


def pme_recip(positions, box, Q, kappa, K, lmax, kind=1, gamma=False, jacobian=None,
              korder=None):
    """recip.py:394-426 (the body of the generated ``pme_recip`` closure).

    kind: 1 (Coulomb, times DIELECTRIC, gamma point dropped) or 6/8/10 (dispersion,
    gamma point kept).  ``korder='reference'`` reproduces the meshgrid(kz,kx,ky)
    permutation of recip.py:340 (only meaningful for K1=K2=K3).
    """
    jacobian = DEFAULTS['jacobian'] if jacobian is None else jacobian
    korder = DEFAULTS['korder'] if korder is None else korder
    K1, K2, K3 = int(K[0]), int(K[1]), int(K[2])
    mesh = spread(positions, box, Q, (K1, K2, K3), lmax, jacobian)
    dt = positions.dtype
    k1 = torch.as_tensor(kpts_int_1d(K1), dtype=dt)
    k2 = torch.as_tensor(kpts_int_1d(K2), dtype=dt)
    k3 = torch.as_tensor(kpts_int_1d(K3), dtype=dt)
    if korder == 'reference':
        g = torch.meshgrid(k3, k1, k2, indexing='xy')                 # :340
        kint = torch.stack([gi.reshape(-1) for gi in g], dim=1)
        Nv = torch.tensor([K1, K2, K3], dtype=dt)
        m = torch.arange(-2, 3, dtype=dt).reshape(5, 1, 1)
        theta_k = torch.prod(torch.sum(bspline(m + 3.0) * torch.cos(2 * math.pi * m * kint[None] / Nv), dim=0), dim=1)
    else:
        g = torch.meshgrid(k1, k2, k3, indexing='ij')
        kint = torch.stack([gi.reshape(-1) for gi in g], dim=1)
        t1, t2, t3 = theta_k_1d(K1, dt), theta_k_1d(K2, dt), theta_k_1d(K3, dt)
        theta_k = (t1[:, None, None] * t2[None, :, None] * t3[None, None, :]).reshape(-1)
    box_inv = torch.linalg.inv(box)
    kpts = 2 * math.pi * kint @ box_inv                               # :360
    ksq = torch.sum(kpts**2, dim=1)
    V = torch.linalg.det(box)                                         # :409
    S = torch.fft.fftn(mesh).reshape(-1)                              # :410
    S2 = (S.real**2 + S.imag**2) / theta_k**2
    if not gamma:
        C = CK[kind](ksq[1:], kappa, V)
        return torch.sum(C * S2[1:]) * DIELECTRIC                     # :413-424
    # erfc(0)*0 terms are finite; sqrt(0) has an infinite slope, keep autograd finite
    ksq_safe = torch.cat([ksq[:1].detach() * 0 + 0.0, ksq[1:]])
    C0 = CK[kind](ksq_safe[:1], kappa, V)
    C = torch.cat([C0, CK[kind](ksq[1:], kappa, V)])
    return torch.sum(C * S2)
