"""The closed-form adjoints the CUDA kernels implement (tools/analytic_proto.py) must
equal the oracle's autograd (== the reference's jax.grad semantics).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import fixtures, pairlist
from oracle import realspace as orc
from oracle import reciprocal as orecip
from oracle.frames import construct_local_frames
from oracle.harmonics import rot_local2global, cart_dipole_to_harm
from tools import analytic_proto as ap

torch.set_num_threads(4)


@pytest.fixture(scope='module')
def small():
    s = fixtures.water1024().carve(0.36)            # 18 A box, ~40 waters
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 6.0)
    rng = np.random.default_rng(3)
    U = rng.normal(0, 0.05, (s.n_atoms, 3)) * (s.pol.numpy() > 0.001)[:, None]
    # make every site anisotropic + polarizable so that all terms are exercised
    Ql = s.Q_local.numpy().copy()
    Ql[:, 1:] += rng.normal(0, 0.05, (s.n_atoms, 8))
    pol = np.abs(rng.normal(0.8, 0.2, s.n_atoms))
    pol[1::7] = 0.0
    th = np.abs(rng.normal(3.0, 1.0, s.n_atoms))
    U = rng.normal(0, 0.05, (s.n_atoms, 3))
    return s, pairs[:n], Ql, U, pol, th


def _t(x, g=True):
    return torch.tensor(np.asarray(x), dtype=torch.float64, requires_grad=g)


def test_bspline_recursion_matches_truncated_power():
    f = np.random.default_rng(0).uniform(0, 1, 50)
    f[0] = 0.0
    w = ap.bspline6_all(f)
    for k in range(6):
        u = torch.tensor(f + k, dtype=torch.float64, requires_grad=True)
        m = orecip.bspline(u)
        np.testing.assert_allclose(w[0, k], m.detach().numpy(), atol=2e-13)
        np.testing.assert_allclose(w[1, k], orecip.bspline_prime(u).detach().numpy(), atol=2e-13)
        np.testing.assert_allclose(w[2, k], orecip.bspline_prime2(u).detach().numpy(), atol=2e-12)
        d3 = torch.autograd.grad(orecip.bspline_prime2(u).sum(), u)[0]
        np.testing.assert_allclose(w[3, k], d3.numpy(), atol=2e-12)


@pytest.mark.parametrize('lpol', [False, True])
def test_pair_real_adjoints(small, lpol):
    s, pairs, Ql, U, pol, th = small
    kappa = 0.45
    pos, box = _t(s.positions), _t(s.box)
    fr = construct_local_frames(s.positions, s.box, s.axis_type, s.axis_indices)
    Qg = _t(rot_local2global(torch.tensor(Ql), fr, 2))
    Uh = _t(cart_dipole_to_harm(torch.tensor(U)))
    tpol, tth = _t(pol), _t(th)
    mS, pS, dS = _t([0.1, 0.3, 0.0, 0.7, 1.0]), _t([0.0, 0.4, 0.0, 1.0, 1.0]), _t([0., 0., 0., 1., 1.])
    E = orc.pme_real(pos, box, pairs, Qg, Uh if lpol else None, tpol if lpol else None, tth if lpol else None,
                     mS, pS if lpol else None, dS if lpol else None, s.covalent_map, kappa, 2, lpol)
    ins = [pos, box, Qg, mS] + ([Uh, tth, tpol, pS] if lpol else [])
    gr = torch.autograd.grad(E, ins, allow_unused=True)
    sidx = orc.pair_scale_index(torch.as_tensor(pairs.astype(np.int64)), s.covalent_map).numpy() % 5
    M = ap.harm_to_cart(Qg.detach().numpy())
    out = ap.pair_real(s.positions.numpy(), np.diag(s.box.numpy()), pairs.astype(np.int64), M, U, pol, th,
                       mS.detach().numpy(), pS.detach().numpy(), sidx, kappa, lpol)
    assert abs(out['E'] - E.item()) < 1e-9 * abs(E.item())
    tol = dict(rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(out['dpos'], gr[0].numpy(), **tol)
    np.testing.assert_allclose(out['dbox'], gr[1].numpy(), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(ap.cart_grad_to_harm(out['G']), gr[2].numpy(), **tol)
    np.testing.assert_allclose(out['dmS'], gr[3].numpy(), **tol)
    if lpol:
        np.testing.assert_allclose(out['F'][:, [2, 0, 1]], gr[4].numpy(), **tol)
        np.testing.assert_allclose(out['dthole'], gr[5].numpy(), **tol)
        np.testing.assert_allclose(out['dpol'], gr[6].numpy(), **tol)
        np.testing.assert_allclose(out['dpS'], gr[7].numpy(), **tol)


@pytest.mark.parametrize('kind', [1, 6, 8, 10])
def test_recip_adjoints(small, kind):
    s, pairs, Ql, U, pol, th = small
    kappa, K = 0.45, (20, 24, 18)
    fr = construct_local_frames(s.positions, s.box, s.axis_type, s.axis_indices)
    Qg0 = rot_local2global(torch.tensor(Ql), fr, 2)
    lmax = 2 if kind == 1 else 0
    if kind != 1:
        Qg0 = torch.cat([Qg0[:, :1].abs() * 10, torch.zeros(s.n_atoms, 8, dtype=torch.float64)], 1)
    pos, box, Qg = _t(s.positions), _t(s.box), _t(Qg0)
    E = orecip.pme_recip(pos, box, Qg[:, :(lmax + 1) ** 2], kappa, K, lmax, kind=kind, gamma=(kind != 1))
    gr = torch.autograd.grad(E, [pos, box, Qg])
    out = ap.recip_all(s.positions.numpy(), s.box.numpy(), K, ap.harm_to_cart(Qg0.numpy()), kappa, kind)
    assert abs(out['E'] - E.item()) < 1e-10 * abs(E.item())
    np.testing.assert_allclose(out['dpos'], gr[0].numpy(), rtol=1e-8, atol=1e-8 * np.abs(gr[0].numpy()).max())
    nh = (lmax + 1) ** 2          # dispersion passes are lmax = 0: only the 'charge' slot is live
    np.testing.assert_allclose(ap.cart_grad_to_harm(out['G'])[:, :nh], gr[2].numpy()[:, :nh], rtol=1e-8,
                               atol=1e-9 * np.abs(gr[2].numpy()).max())
    np.testing.assert_allclose(out['dbox'], gr[1].numpy(), rtol=1e-7, atol=1e-8 * np.abs(gr[1].numpy()).max())


def test_frames_and_self_adjoints(small):
    s, pairs, Ql, U, pol, th = small
    pos, box, tQl = _t(s.positions), _t(s.box), _t(Ql)
    fr = construct_local_frames(pos, box, s.axis_type, s.axis_indices)
    Qg = rot_local2global(tQl, fr, 2)
    rng = np.random.default_rng(5)
    Gh = rng.normal(0, 1, (s.n_atoms, 9))
    L = torch.sum(Qg * torch.tensor(Gh))
    gr = torch.autograd.grad(L, [pos, box, tQl])
    # Cartesian gradient equivalent to Gh: dL/dM = pinv-chain; use the adjoint identity
    # L = sum Gh . Q(M)  with Q(M) the linear inverse of harm_to_cart
    Minv = np.linalg.pinv(np.stack([ap.harm_to_cart(np.eye(9)[k:k + 1])[0] for k in range(9)]).T)  # (9,10)
    G = Gh @ Minv
    out = ap.frames_fwd_bwd(s.positions.numpy(), np.diag(s.box.numpy()), s.axis_type, s.axis_indices, Ql, G)
    np.testing.assert_allclose(out['M'], ap.harm_to_cart(Qg.detach().numpy()), rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(out['R'], fr.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out['dQ'], gr[2].numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out['dpos'], gr[0].numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(out['dbox'], gr[1].numpy(), rtol=1e-8, atol=1e-8)
    # self + penalty
    tU, tpol = _t(U), _t(pol)
    Qg2 = _t(Qg.detach())
    Uh = cart_dipole_to_harm(tU)
    Qtot = torch.cat([Qg2[:, :1], Qg2[:, 1:4] + Uh, Qg2[:, 4:]], 1)
    Es = orc.pme_self(Qtot, 0.45, 2) + orc.pol_penalty(Uh, tpol)
    g2 = torch.autograd.grad(Es, [Qg2, tU, tpol])
    o2 = ap.self_terms(ap.harm_to_cart(Qg.detach().numpy()), U, pol, 0.45, True)
    assert abs(o2['E'] - Es.item()) < 1e-10 * abs(Es.item())
    np.testing.assert_allclose(ap.cart_grad_to_harm(o2['G']), g2[0].numpy(), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(o2['F'], g2[1].numpy(), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(o2['dpol'], g2[2].numpy(), rtol=1e-10, atol=1e-9)
