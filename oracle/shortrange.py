"""Oracle (test infrastructure): generic pair-sum driver and the Tang-Toennies kernel.
Restates admp/pairwise.py:45-113."""
import torch

from .frames import pbc_shift
from .realspace import filter_pairs, pair_scale_index


def TT_damping_qq_c6_kernel(dr, m, ai, aj, bi, bj, qi, qj, ci, cj):
    """admp/pairwise.py:94-113 (constants 2625.5 Ha->kJ/mol, 1.889726878 A->bohr)."""
    a = torch.sqrt(ai * aj)
    b = torch.sqrt(bi * bj)
    c = ci * cj
    q = qi * qj
    br = b * dr * 1.889726878
    poly, term = torch.ones_like(br), torch.ones_like(br)
    for k in range(1, 7):
        term = term * br / k
        poly = poly + term
    e = torch.exp(-br)
    f = 2625.5 * a * e + (-2625.5) * e * (1 + br) * q / br + e * poly * c / dr**6
    return f * m


def generate_pairwise_interaction(pair_int_kernel, covalent_map, static_args=None):
    """admp/pairwise.py:45-91."""
    def pair_int(positions, box, pairs, mScales, *atomic_params):
        pr = filter_pairs(pairs)
        i, j = pr[:, 0], pr[:, 1]
        m = mScales[pair_scale_index(pr, covalent_map)]
        dr = torch.linalg.norm(pbc_shift(positions[i] - positions[j], box), dim=1)
        pp = []
        for prm in atomic_params:
            pp += [prm[i], prm[j]]
        return torch.sum(pair_int_kernel(dr, m, *pp))
    return pair_int
