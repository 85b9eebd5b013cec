"""Oracle (test infrastructure): PBC shift, per-site local frames, pair QI frame.

Restates admp/spatial.py.
"""
import numpy as np
import torch

ZTHENX, BISECTOR, ZBISECT, THREEFOLD, ZONLY, NOAXIS = 0, 1, 2, 3, 4, 5   # spatial.py:59-64


def pbc_shift(dr, box, box_inv=None):
    """admp/spatial.py:13-32.  Row-vector convention, box rows = lattice vectors.

    ``ds - floor(ds + 0.5)`` => a component exactly at +L/2 maps to -L/2 (A14).
    """
    if box_inv is None:
        box_inv = torch.linalg.inv(box)
    ds = dr @ box_inv
    ds = ds - torch.floor(ds + 0.5)
    return ds @ box


def _unit(v):
    return v / torch.linalg.norm(v, dim=1, keepdim=True)        # spatial.py:36-41


def construct_local_frames(positions, box, axis_types, axis_indices):
    """admp/spatial.py:76-142.  Returns (n,3,3) frames with rows (x, y, z).

    Sites of type NOAXIS get the identity frame.  (The reference would index atom
    -1 for them, i.e. build a frame towards the last atom; such sites carry no
    anisotropic moments in any shipped input, so this is a documented divergence
    with no observable effect there.)
    """
    axis_types = np.asarray(axis_types)
    ai = torch.as_tensor(np.asarray(axis_indices), dtype=torch.long)
    n = positions.shape[0]
    box_inv = torch.linalg.inv(box)
    is_zonly = torch.as_tensor(axis_types == ZONLY)
    is_bis = torch.as_tensor(axis_types == BISECTOR)
    is_zbis = torch.as_tensor(axis_types == ZBISECT)
    is_three = torch.as_tensor(axis_types == THREEFOLD)
    # anchors that a site's axis type never reads are pointed at a neighbour so the
    # masked-out branch stays finite (keeps autograd free of 0*nan)
    other = (torch.arange(n) + 1) % n
    is_none = torch.as_tensor(axis_types == NOAXIS)
    zat = torch.where(is_none, other, ai[:, 0])
    is_zonly = torch.logical_or(is_zonly, is_none)
    xat = torch.where(is_zonly, other, ai[:, 1])
    yat = torch.where(torch.logical_or(is_zbis, is_three), ai[:, 2], other)

    vz = _unit(pbc_shift(positions[zat] - positions, box, box_inv))          # :98-99
    # Z-only: x = (1 - round|z_x|, round|z_x|, 0)                              :103-105
    xz0 = torch.round(torch.abs(vz[:, 0]))
    vx_zonly = torch.stack([1.0 - xz0, xz0, torch.zeros_like(xz0)], dim=1)
    # others: normalised vector to the x anchor                               :107-110
    vx_other = pbc_shift(positions[xat] - positions, box, box_inv)
    vx = torch.where(is_zonly[:, None], vx_zonly, _unit(vx_other))

    if bool(is_bis.any()):                                                    # :112-114
        vz = torch.where(is_bis[:, None], _unit(vz + vx), vz)
    vy = torch.zeros_like(vz)
    if bool(is_zbis.any()):                                                   # :116-121
        vyz = _unit(pbc_shift(positions[yat] - positions, box, box_inv))
        vx = torch.where(is_zbis[:, None], _unit(vx + vyz), vx)
    if bool(is_three.any()):                                                  # :123-135
        vy3 = _unit(pbc_shift(positions[yat] - positions, box, box_inv))
        vz = torch.where(is_three[:, None], _unit(vz + vx + vy3), vz)

    proj = torch.sum(vx * vz, dim=1, keepdim=True)                            # :138-139
    vx = _unit(vx - vz * proj)
    vy = torch.linalg.cross(vz, vx, dim=1)                                    # :141
    frames = torch.stack([vx, vy, vz], dim=1)
    if bool(is_none.any()):
        eye = torch.eye(3, dtype=frames.dtype).expand(n, 3, 3)
        frames = torch.where(is_none[:, None, None], eye, frames)
    return frames


def build_quasi_internal(r1, r2, dr, norm_dr):
    """admp/spatial.py:149-178.  z = dr/|dr| (points from r2 to r1), x by
    Gram-Schmidt of z+(1,0,0) unless the RAW r1,r2 agree in y and z (A13)."""
    vz = dr / norm_dr[:, None]
    use_x = torch.logical_or(r1[:, 1] != r2[:, 1], r1[:, 2] != r2[:, 2])
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=dr.dtype)
    ey = torch.tensor([0.0, 1.0, 0.0], dtype=dr.dtype)
    vx = torch.where(use_x[:, None], vz + ex, vz + ey)
    vx = vx - vz * torch.sum(vz * vx, dim=1, keepdim=True)
    vx = vx / torch.linalg.norm(vx, dim=1, keepdim=True)
    vy = torch.linalg.cross(vz, vx, dim=1)
    return torch.stack([vx, vy, vz], dim=1)
