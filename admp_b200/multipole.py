"""Multipole conversions and rotations - drop-in surface of admp/multipole.py.

Component order: 00, 10, 11c, 11s, 20, 21c, 21s, 22c, 22s.  ``convert_cart2harm`` is a
fixed linear map applied once at set-up time (host-side parameter preparation, as in the
reference's scripts); the rotations run on the device (admp_rotate).
"""
import numpy as np
import torch

from . import _lib
from ._ctx import Context, to_dev

rt3 = 1.73205080757          # admp/multipole.py:14
inv_rt3 = 1.0 / rt3

# admp/multipole.py:17-33
C1_h2c = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
C1_c2h = C1_h2c.T
C2_c2h = np.array([[0, 0, 1, 0, 0, 0],
                   [0, 0, 0, 0, 2 * inv_rt3, 0],
                   [0, 0, 0, 0, 0, 2 * inv_rt3],
                   [inv_rt3, -inv_rt3, 0, 0, 0, 0],
                   [0, 0, 0, 2 * inv_rt3, 0, 0]], dtype=np.float64)
C2_h2c = np.array([[-0.5, 0, 0, rt3 / 2, 0],
                   [-0.5, 0, 0, -rt3 / 2, 0],
                   [1, 0, 0, 0, 0],
                   [0, 0, 0, 0, rt3 / 2],
                   [0, rt3 / 2, 0, 0, 0],
                   [0, 0, rt3 / 2, 0, 0]], dtype=np.float64)

_ctx_cache = {}


def _ctx():
    from . import settings
    key = (settings.PRECISION, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    if key not in _ctx_cache:
        _ctx_cache[key] = Context()
    return _ctx_cache[key]


def convert_cart2harm(Theta, lmax):
    '''
    admp/multipole.py:36-77: Cartesian moments (n, 10) ``c0, dX, dY, dZ, qXX, qYY, qZZ, qXY, qXZ, qYZ``
    -> harmonic (n, (lmax+1)^2).  Accepts numpy arrays or tensors; returns the same kind.
    '''
    if lmax > 2:
        raise NotImplementedError('l > 2 (beyond quadrupole) not supported')
    is_t = isinstance(Theta, torch.Tensor)
    T = Theta if is_t else torch.as_tensor(np.asarray(Theta, dtype=np.float64))
    cols = [T[:, 0:1]]
    if lmax >= 1:
        cols.append(T[:, 1:4] @ torch.as_tensor(C1_c2h.T, dtype=T.dtype, device=T.device))
    if lmax >= 2:
        cols.append(T[:, 4:10] @ torch.as_tensor(C2_c2h.T, dtype=T.dtype, device=T.device))
    Q = torch.cat(cols, dim=1)
    return Q if is_t else Q.numpy()


def _rotate(Q, localframes, lmax, to_local):
    if lmax > 2:
        raise NotImplementedError('l > 2 (beyond quadrupole) not supported')
    cx = _ctx()
    Q = to_dev(Q, cx.dtype, cx.device).detach()
    R = to_dev(localframes, cx.dtype, cx.device).detach().reshape(-1, 9)
    nh = (lmax + 1) ** 2
    if Q.shape != (R.shape[0], nh):
        raise ValueError('Q must be (n, %d) and localframes (n, 3, 3)' % nh)
    out = torch.empty_like(Q)
    _lib.check(cx.lib.admp_rotate(cx.handle, _lib.stream_ptr(), int(Q.shape[0]), int(lmax), int(to_local),
                                  _lib.ptr(Q), _lib.ptr(R), _lib.ptr(out)))
    return out


def rot_global2local(Q_gh, localframes, lmax=2):
    '''admp/multipole.py:92-179: rotate harmonic moments from the global to the local frame.'''
    return _rotate(Q_gh, localframes, lmax, 1)


def rot_local2global(Q_lh, localframes, lmax=2):
    '''admp/multipole.py:183-201: rot_global2local with the transposed frame.'''
    return _rotate(Q_lh, localframes, lmax, 0)


def rot_ind_global2local(U_g, localframes):
    '''admp/multipole.py:80-89: dipole-only rotation, harmonic (z,x,y) component order.'''
    U_g = torch.as_tensor(np.asarray(U_g)) if not isinstance(U_g, torch.Tensor) else U_g
    Q = torch.cat([torch.zeros_like(U_g[:, :1]), U_g], dim=1)
    return _rotate(Q, localframes, 1, 1)[:, 1:4]
