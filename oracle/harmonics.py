"""Oracle (test infrastructure): real-spherical-harmonic multipole algebra.

Restates admp/multipole.py.  Component order everywhere:
``00, 10, 11c, 11s, 20, 21c, 21s, 22c, 22s``.
"""
import torch

# admp/multipole.py:14 - the reference uses this 12-digit literal, not sqrt(3)
RT3 = 1.73205080757
INV_RT3 = 1.0 / RT3

# harmonic dipole order is (z, x, y): admp/multipole.py:17-20
_ZXY = [2, 0, 1]


def cart2harm(theta, lmax=2):
    """admp/multipole.py:36-77 (convert_cart2harm).

    theta: (n, 10) Cartesian moments ``c0, dX, dY, dZ, qXX, qYY, qZZ, qXY, qXZ, qYZ``
    returns (n, (lmax+1)^2) harmonic moments.
    """
    if lmax > 2:
        raise NotImplementedError('l > 2 (beyond quadrupole) not supported')
    theta = torch.as_tensor(theta, dtype=torch.float64)
    cols = [theta[:, 0]]
    if lmax >= 1:
        # C1_c2h . (dX, dY, dZ) = (dZ, dX, dY)      multipole.py:17-20,63-64
        cols += [theta[:, 3], theta[:, 1], theta[:, 2]]
    if lmax >= 2:
        xx, yy, zz, xy, xz, yz = (theta[:, 4 + k] for k in range(6))
        # rows of C2_c2h, multipole.py:22-26
        cols += [zz, 2 * INV_RT3 * xz, 2 * INV_RT3 * yz,
                 INV_RT3 * xx - INV_RT3 * yy, 2 * INV_RT3 * xy]
    return torch.stack(cols, dim=1)


def quad_rotation_matrix(R):
    """5x5 matrix D with Q2_local = D @ Q2_global for frames R (n,3,3), rows = axes.

    admp/multipole.py:124-171: the polynomial-in-frame-entries matrix ``C2_gl``;
    D[:, j, k] is the reference's ``C2_gl_jk``.
    """
    xx, xy, xz = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    yx, yy, yz = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    zx, zy, zz = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    r = RT3
    rows = [
        [(3 * zz**2 - 1) / 2, r * zx * zz, r * zy * zz,
         (r * (-2 * zy**2 - zz**2 + 1)) / 2, r * zx * zy],
        [r * xz * zz, 2 * xx * zz - yy, yx + 2 * xy * zz,
         -2 * xy * zy - xz * zz, xx * zy + zx * xy],
        [r * yz * zz, 2 * yx * zz + xy, -xx + 2 * yy * zz,
         -2 * yy * zy - yz * zz, yx * zy + zx * yy],
        [r * (-2 * yz**2 - zz**2 + 1) / 2, -2 * yx * yz - zx * zz, -2 * yy * yz - zy * zz,
         (4 * yy**2 + 2 * zy**2 + 2 * yz**2 + zz**2 - 3) / 2, -2 * yx * yy - zx * zy],
        [r * xz * yz, xx * yz + yx * xz, xy * yz + yy * xz,
         -2 * xy * yy - xz * yz, xx * yy + yx * xy],
    ]
    return torch.stack([torch.stack(row, dim=1) for row in rows], dim=1)


def rot_global2local(Q, R, lmax=2):
    """admp/multipole.py:92-179.  Q (n,(lmax+1)^2), R (n,3,3) -> rotated Q."""
    if lmax > 2:
        raise NotImplementedError('l > 2 (beyond quadrupole) not supported')
    out = [Q[:, 0:1]]
    if lmax >= 1:
        R1 = R[:, _ZXY][:, :, _ZXY]                       # multipole.py:118-120
        out.append(torch.einsum('nij,nj->ni', R1, Q[:, 1:4]))
    if lmax >= 2:
        D = quad_rotation_matrix(R)
        out.append(torch.einsum('njk,nk->nj', D, Q[:, 4:9]))   # multipole.py:171
    return torch.cat(out, dim=1)


def rot_local2global(Q, R, lmax=2):
    """admp/multipole.py:183-201: global2local with the transposed frame."""
    return rot_global2local(Q, R.transpose(-2, -1), lmax)


def rot_ind_global2local(U, R):
    """admp/multipole.py:80-89: dipole-only rotation (harmonic z,x,y order)."""
    R1 = R[:, _ZXY][:, :, _ZXY]
    return torch.einsum('nij,nj->ni', R1, U)


def cart_dipole_to_harm(U):
    """C1_c2h applied to Cartesian (x,y,z) dipoles -> (z,x,y): admp/pme.py:235."""
    return U[:, _ZXY]
