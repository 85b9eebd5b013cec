"""Self-checks of the oracle that need no reference output: finite differences,
Ewald-parameter invariance, translation invariance, replica invariance. CPU only, small systems."""
import numpy as np
import pytest
import torch

from oracle import fixtures, pairlist
from oracle import realspace as orc
from oracle.dispersion import energy_disp_pme

torch.set_num_threads(4)


@pytest.fixture(scope='module')
def sysm():
    s = fixtures.lattice_water(3, 3.3, seed=2)        # 27 waters, 9.9 A box
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.9)
    rng = np.random.default_rng(0)
    U = torch.tensor(rng.normal(0, 0.03, (s.n_atoms, 3))) * (s.pol > 0.001)[:, None]
    return s, pairs, U


def _E(s, pairs, U, pos, box, kappa=0.55, K=24):
    return orc.energy_pme(pos, box, pairs, s.Q_local, U, s.pol, s.tholes, s.mScales, s.pScales, s.dScales, s.covalent_map,
                          s.axis_type, s.axis_indices, kappa, K, K, K, 2, True)


def test_forces_match_finite_differences(sysm):
    s, pairs, U = sysm
    pos = s.positions.clone().requires_grad_(True)
    E = _E(s, pairs, U, pos, s.box)
    g = torch.autograd.grad(E, pos)[0]
    h = 1e-5
    for (a, c) in [(0, 0), (4, 2), (11, 1)]:
        p1, p2 = s.positions.clone(), s.positions.clone()
        p1[a, c] += h
        p2[a, c] -= h
        fd = (_E(s, pairs, U, p1, s.box) - _E(s, pairs, U, p2, s.box)) / (2 * h)
        assert abs(fd.item() - g[a, c].item()) < 1e-5 * max(1.0, abs(g[a, c].item()))


def test_virial_diagonal_matches_finite_differences(sysm):
    s, pairs, U = sysm
    box = s.box.clone().requires_grad_(True)
    g = torch.autograd.grad(_E(s, pairs, U, s.positions, box), box)[0]
    h = 1e-5
    b1, b2 = s.box.clone(), s.box.clone()
    b1[1, 1] += h
    b2[1, 1] -= h
    fd = (_E(s, pairs, U, s.positions, b1) - _E(s, pairs, U, s.positions, b2)) / (2 * h)
    assert abs(fd.item() - g[1, 1].item()) < 1e-5 * abs(g[1, 1].item())


def test_energy_is_independent_of_ewald_parameters(sysm):
    s, _, U = sysm
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.94)
    e1 = _E(s, pairs, U, s.positions, s.box, kappa=0.80, K=40).item()
    e2 = _E(s, pairs, U, s.positions, s.box, kappa=0.90, K=48).item()
    assert abs(e1 - e2) < 2e-4 * abs(e1)


def test_translation_invariance_and_zero_net_force(sysm):
    s, pairs, U = sysm
    pos = s.positions.clone().requires_grad_(True)
    E = _E(s, pairs, U, pos, s.box)
    g = torch.autograd.grad(E, pos)[0]
    assert g.sum(0).abs().max().item() < 1e-3 * g.abs().max().item()           # up to PME aliasing
    E2 = _E(s, pairs, U, s.positions + torch.tensor([0.37, -1.2, 2.9]), s.box)
    assert abs(E2.item() - E.item()) < 1e-4 * abs(E.item())


def test_replica_invariance(sysm):
    s, pairs, U = sysm
    r = s.replicate(2, 1, 1)
    pr, _ = pairlist.build_pairs(r.positions.numpy(), r.box.numpy(), 4.9)
    E1 = _E(s, pairs, U, s.positions, s.box, kappa=0.55, K=24)
    E2 = orc.energy_pme(r.positions, r.box, pr, r.Q_local, U.repeat(2, 1), r.pol, r.tholes, r.mScales, r.pScales, r.dScales,
                        r.covalent_map, r.axis_type, r.axis_indices, 0.55, 48, 24, 24, 2, True)
    assert abs(E2.item() - 2 * E1.item()) < 1e-9 * abs(E1.item())


def test_dispersion_matches_direct_lattice_sum(sysm):
    """C6 part of the dispersion PME against a brute-force minimum-image + image-shell sum."""
    s, _, _ = sysm
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.94)
    m1 = torch.ones(5, dtype=torch.float64)
    E = energy_disp_pme(s.positions, s.box, pairs, s.c_list, m1, s.covalent_map, 0.9, 40, 40, 40, 6).item()
    pos, L, c6 = s.positions.numpy(), s.box[0, 0].item(), s.c_list[:, 0].numpy()
    tot = 0.0
    R = 4
    sh = np.stack(np.meshgrid(*[np.arange(-R, R + 1)] * 3, indexing='ij'), -1).reshape(-1, 3) * L
    for i in range(s.n_atoms):
        d = pos[None, :, :] - pos[i][None, None, :] + sh[:, None, :]
        r2 = (d ** 2).sum(-1)
        r2[np.all(sh == 0, axis=1), i] = np.inf
        tot += 0.5 * np.sum(c6[i] * c6[None, :] / r2 ** 3)
    assert abs(E - tot) < 2e-3 * abs(tot)           # shell truncation of the brute-force sum


def test_c8_and_c10_dispersion_match_direct_lattice_sums(sysm):
    """The r^-8 and r^-10 parts of the dispersion PME (admp/disp_pme.py:80-279, Ck_8 / Ck_10 of admp/recip.py:445-462),
    isolated as differences of pmax = 10, 8, 6 evaluations, against brute-force image sums; these converge absolutely
    and fast (shell truncation at 4 L < 1e-6), so the comparison is tight."""
    s, _, _ = sysm
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.94)
    m1 = torch.ones(5, dtype=torch.float64)
    E = {p: energy_disp_pme(s.positions, s.box, pairs, s.c_list, m1, s.covalent_map, 0.9, 40, 40, 40, p).item() for p in (6, 8, 10)}
    pos, L = s.positions.numpy(), s.box[0, 0].item()
    R = 4
    sh = np.stack(np.meshgrid(*[np.arange(-R, R + 1)] * 3, indexing='ij'), -1).reshape(-1, 3) * L
    home = np.all(sh == 0, axis=1)
    for col, p, part in ((1, 8, E[8] - E[6]), (2, 10, E[10] - E[8])):
        c = s.c_list[:, col].numpy()
        tot = 0.0
        for i in range(s.n_atoms):
            d = pos[None, :, :] - pos[i][None, None, :] + sh[:, None, :]
            r2 = (d ** 2).sum(-1)
            r2[home, i] = np.inf
            tot += 0.5 * np.sum(c[i] * c[None, :] / r2 ** (p // 2))
        assert abs(part - tot) < 1e-5 * abs(tot), (p, part, tot)
