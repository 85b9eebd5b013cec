"""Spatial helpers - drop-in surface of admp/spatial.py.

``generate_construct_local_frames`` is the device kernel used by the calculators.
``pbc_shift`` / ``v_pbc_shift`` / ``build_quasi_internal`` are small helper functions the
reference's tests exercise; they are not on the device hot path here (the pair kernel works
with rotational invariants instead of a quasi-internal frame, and applies the minimum image
inline), so they are provided as plain tensor arithmetic for API completeness.
"""
import numpy as np
import torch

from . import _lib
from ._ctx import Context, to_dev


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))


def pbc_shift(drvecs, box, box_inv):
    '''admp/spatial.py:13-32: ds = dr.box_inv ; ds -= floor(ds + 0.5) ; return ds.box'''
    dr, box, box_inv = _t(drvecs).to(torch.float64), _t(box).to(torch.float64), _t(box_inv).to(torch.float64)
    ds = dr @ box_inv
    ds = ds - torch.floor(ds + 0.5)
    return ds @ box


def normalize(matrix, axis=1, ord=2):
    '''admp/spatial.py:36-41: normalise a matrix along one dimension'''
    m = _t(matrix)
    return m / torch.linalg.norm(m, ord=ord, dim=axis, keepdim=True)


v_pbc_shift = pbc_shift          # admp/spatial.py:34 (vmap over rows; the matrix form is already row-wise)


def build_quasi_internal(r1, r2, dr, norm_dr):
    '''admp/spatial.py:149-178: pair frames with z = dr/|dr| and x from Gram-Schmidt of z+(1,0,0)
    (or z+(0,1,0) when the raw r1, r2 agree in y and z).  Rows are (x, y, z).'''
    r1, r2, dr, nrm = (_t(x).to(torch.float64) for x in (r1, r2, dr, norm_dr))
    vz = dr / nrm[:, None]
    use_x = torch.logical_or(r1[:, 1] != r2[:, 1], r1[:, 2] != r2[:, 2])
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=dr.device)
    ey = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64, device=dr.device)
    vx = torch.where(use_x[:, None], vz + ex, vz + ey)
    vx = vx - vz * torch.sum(vz * vx, dim=1, keepdim=True)
    vx = vx / torch.linalg.norm(vx, dim=1, keepdim=True)
    return torch.stack([vx, torch.linalg.cross(vz, vx, dim=1), vz], dim=1)


def generate_construct_local_frames(axis_types, axis_indices):
    """admp/spatial.py:44-147: returns ``construct_local_frames(positions, box) -> (n, 3, 3)`` with
    rows (x, y, z); axis types ZThenX=0, Bisector=1, ZBisect=2, ThreeFold=3, Zonly=4, NoAxisType=5."""
    axis_types = np.asarray(axis_types)
    axis_indices = np.asarray(axis_indices)
    n = axis_types.shape[0]
    cx = Context()
    cx.set_topology(n, axis_types, axis_indices, None)
    # lmax = 2 so that frames are built; the mesh size is irrelevant for this stage
    cx.set_pme(1.0, 6, 6, 6, 2)

    def construct_local_frames(positions, box):
        positions = to_dev(positions, cx.dtype, cx.device).detach()
        box = to_dev(box, cx.dtype, cx.device).detach()
        if positions.shape != (n, 3):
            raise ValueError('positions must be (%d, 3)' % n)
        fr = torch.empty((n, 3, 3), dtype=cx.dtype, device=cx.device)
        dummy = torch.zeros((n, 9), dtype=cx.dtype, device=cx.device)
        _lib.check(cx.lib.admp_frames_fwd(cx.handle, _lib.stream_ptr(), _lib.ptr(positions), _lib.ptr(box),
                                          _lib.ptr(dummy), None, None, _lib.ptr(fr)))
        return fr

    construct_local_frames._ctx = cx
    construct_local_frames.axis_types = axis_types            # read back by admp_b200.pme.energy_pme
    construct_local_frames.axis_indices = axis_indices
    return construct_local_frames
