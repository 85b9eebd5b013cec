"""Oracle (test infrastructure): C6/C8/C10 dispersion PME.  Restates admp/disp_pme.py."""
import numpy as np
import torch

from .frames import pbc_shift
from .realspace import filter_pairs, pair_scale_index, setup_ewald_parameters
from .reciprocal import pme_recip


def g_p(x2, pmax):
    """admp/disp_pme.py:219-251: g_p = exp(-x^2) sum_{k<p/2} x^{2k}/k!  for p = 6, 8, 10."""
    x4 = x2 * x2
    e = torch.exp(-x2)
    g = [1 + x2 + 0.5 * x4]
    if pmax >= 8:
        g.append(g[0] + x4 * x2 / 6)
    if pmax >= 10:
        g.append(g[1] + x4 * x4 / 24)
    return [gi * e for gi in g]


def disp_pme_real(positions, box, pairs, c_list, mScales, covalent_map, kappa, pmax):
    """admp/disp_pme.py:126-216 (driver + per-pair kernel)."""
    pairs = filter_pairs(pairs)
    i, j = pairs[:, 0], pairs[:, 1]
    m = mScales[pair_scale_index(pairs, covalent_map)]
    dr = pbc_shift(positions[i] - positions[j], box)
    dr2 = torch.sum(dr * dr, dim=1)
    g = g_p(kappa * kappa * dr2, pmax)
    ci, cj = c_list[i], c_list[j]
    dr6 = dr2 * dr2 * dr2
    ene = (m + g[0] - 1) * ci[:, 0] * cj[:, 0] / dr6
    if pmax >= 8:
        dr8 = dr6 * dr2
        ene = ene + (m + g[1] - 1) * ci[:, 1] * cj[:, 1] / dr8
    if pmax >= 10:
        ene = ene + (m + g[2] - 1) * ci[:, 2] * cj[:, 2] / (dr8 * dr2)
    return torch.sum(ene)


def disp_pme_self(c_list, kappa, pmax):
    """admp/disp_pme.py:254-279."""
    E = -kappa**6 / 12 * torch.sum(c_list[:, 0] ** 2)
    if pmax >= 8:
        E = E - kappa**8 / 48 * torch.sum(c_list[:, 1] ** 2)
    if pmax >= 10:
        E = E - kappa**10 / 240 * torch.sum(c_list[:, 2] ** 2)
    return E


def energy_disp_pme(positions, box, pairs, c_list, mScales, covalent_map,
                    kappa, K1, K2, K3, pmax, parts=None):
    """admp/disp_pme.py:80-123: three lmax=0 reciprocal passes, gamma point kept."""
    e_real = disp_pme_real(positions, box, pairs, c_list, mScales, covalent_map, kappa, pmax)
    K = (K1, K2, K3)
    e_recip = pme_recip(positions, box, c_list[:, 0:1], kappa, K, 0, kind=6, gamma=True)
    if pmax >= 8:
        e_recip = e_recip + pme_recip(positions, box, c_list[:, 1:2], kappa, K, 0, kind=8, gamma=True)
    if pmax >= 10:
        e_recip = e_recip + pme_recip(positions, box, c_list[:, 2:3], kappa, K, 0, kind=10, gamma=True)
    e_self = disp_pme_self(c_list, kappa, pmax)
    if parts is not None:
        parts.update(real=e_real.detach(), recip=e_recip.detach(), self=e_self.detach())
    return e_real + e_recip + e_self


class OracleDispPmeForce:
    """Mirror of ``ADMPDispPmeForce`` (admp/disp_pme.py:20-77)."""

    def __init__(self, box, covalent_map, rc, ethresh, pmax):
        self.covalent_map, self.rc, self.ethresh, self.pmax = covalent_map, rc, ethresh, pmax
        self.kappa, self.K1, self.K2, self.K3 = setup_ewald_parameters(rc, ethresh, box)
        self.pme_order = 6

    def update_env(self, attr, val):
        setattr(self, attr, val)

    def get_energy(self, positions, box, pairs, c_list, mScales, parts=None):
        return energy_disp_pme(positions, box, pairs, c_list, mScales, self.covalent_map,
                               self.kappa, self.K1, self.K2, self.K3, self.pmax, parts)
