"""GPU parity tests proper: the CUDA path (through the C ABI of libadmp_b200.so, driven by the
reference-shaped Python surface) against the CPU oracle on the same seeded inputs.

Tolerances (north star): energies and gradients within 1e-6 relative in double precision,
1e-4 relative in single; neighbour pair sets bit-exact; SCF iteration counts identical.
"Relative" for arrays means relative to the largest magnitude of the reference array.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fixtures, pairlist                                   # noqa: E402
from oracle import realspace as orc                                     # noqa: E402
from oracle import reciprocal as orecip                                 # noqa: E402
from oracle.dispersion import OracleDispPmeForce                        # noqa: E402
from oracle.frames import construct_local_frames                        # noqa: E402
from oracle.harmonics import rot_local2global, rot_global2local         # noqa: E402
from oracle.shortrange import generate_pairwise_interaction as o_pairwise, TT_damping_qq_c6_kernel as o_tt  # noqa: E402

RTOL = 1e-6


def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def _t(x, g=True):
    return torch.tensor(np.asarray(x), dtype=torch.float64, requires_grad=g)


@pytest.fixture(scope='module')
def carved():
    s = fixtures.water1024().carve(0.5)        # 99 waters, 25 A box, gas-like
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 6.0)
    return s, pairs


@pytest.fixture(scope='module')
def lattice():
    s = fixtures.lattice_water(4, 3.15, seed=5)  # 64 waters, 12.6 A box, liquid-like (SCF converges)
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 5.0)
    return s, pairs


def _perturbed(s, seed=3):
    rng = np.random.default_rng(seed)
    Ql = s.Q_local.numpy().copy()
    Ql[:, 1:] += rng.normal(0, 0.05, (s.n_atoms, 8))
    pol = np.abs(rng.normal(0.8, 0.2, s.n_atoms))
    pol[1::7] = 0.0
    th = np.abs(rng.normal(3.0, 1.0, s.n_atoms))
    U = rng.normal(0, 0.05, (s.n_atoms, 3))
    return Ql, U, pol, th


# ------------------------------------------------------------------------------------------ helpers / frames
def test_library_loaded_and_graph_scf_available():
    from admp_b200 import _lib
    lib = _lib.load()
    assert lib.admp_version() >= 100
    assert torch.cuda.get_device_capability(0)[0] >= 10, 'built for sm_100a only'


def test_frames_and_rotation_match_reference_literals():
    """tests/test_sptial.py:70-142 and tests/test_multipole.py:83-189 of the reference."""
    from admp_b200.spatial import generate_construct_local_frames
    from admp_b200.multipole import rot_local2global as g_l2g, rot_global2local as g_g2l
    s = fixtures.water2()
    fn = generate_construct_local_frames(s.axis_type, s.axis_indices)
    fr = fn(s.positions, s.box).cpu()
    expected0 = np.array([[-0.96165454, -0.17201543, 0.21361469], [0.10460715, -0.95003253, -0.29410106],
                          [0.2535308, -0.26047802, 0.9315972]])
    expected4 = np.array([[-0.5616504, -0.8264594, 0.03890521], [-0.22607785, 0.10806668, -0.9680963],
                          [0.79588807, -0.5525272, -0.24753988]])
    np.testing.assert_allclose(fr[0].numpy(), expected0, rtol=1e-6, atol=2e-6)   # float32-era literals
    np.testing.assert_allclose(fr[4].numpy(), expected4, rtol=1e-6, atol=2e-6)
    ofr = construct_local_frames(s.positions, s.box, s.axis_type, s.axis_indices)
    assert rel(fr, ofr) < 1e-12
    Qg = g_l2g(s.Q_local, fr, 2).cpu()
    Qg_lit0 = np.array([-1.0614, -0.22052474, -0.06001501, 0.06165953, -0.05764905, -0.03114612, 0.02651503,
                        0.01010778, 0.01109856])
    np.testing.assert_allclose(Qg[0].numpy(), Qg_lit0, rtol=1e-6, atol=1e-6)
    assert rel(Qg, rot_local2global(s.Q_local, ofr, 2)) < 1e-10
    back = g_g2l(Qg, fr, 2).cpu()
    np.testing.assert_allclose(back.numpy(), s.Q_local.numpy(), rtol=1e-6, atol=1e-6)
    assert rel(back, rot_global2local(rot_local2global(s.Q_local, ofr, 2), ofr, 2)) < 1e-10


def test_all_axis_types_frames_and_adjoint():
    """ZThenX, Bisector, ZBisect, ThreeFold, Zonly, NoAxisType (admp/spatial.py:59-64): frames and
    the torque -> position adjoint against oracle autograd."""
    from admp_b200.pme import ADMPPmeForce
    rng = np.random.default_rng(9)
    n = 24
    pos = rng.uniform(0, 12.0, (n, 3))
    box = np.diag([12.0, 13.0, 14.0])
    types = np.array([0, 1, 2, 3, 4, 5] * 4)
    ai = np.stack([(np.arange(n) + 1) % n, (np.arange(n) + 2) % n, (np.arange(n) + 3) % n], 1)
    Ql = rng.normal(0, 0.3, (n, 9))
    cov = np.zeros((n, n), dtype=int)
    pairs, _ = pairlist.build_pairs(pos, box, 5.0)
    calc = ADMPPmeForce(box, types, ai, cov, 5.0, 1e-4, 2)
    fr = calc.construct_local_frames(pos, box).cpu()
    tp, tb = _t(pos), _t(box)
    ofr = construct_local_frames(tp, tb, types, ai)
    assert rel(fr, ofr) < 1e-12
    mS = _t([0.0, 0.0, 0.0, 1.0, 1.0], False)
    tQ = _t(Ql)
    Eo = orc.energy_pme(tp, tb, pairs, tQ, None, None, None, mS, None, None, cov, types, ai,
                        calc.kappa, calc.K1, calc.K2, calc.K3, 2, False)
    go = torch.autograd.grad(Eo, [tp, tb, tQ])
    p, b, q = (torch.tensor(x, device='cuda', requires_grad=True) for x in (pos, box, Ql))
    E = calc.get_energy(p, b, pairs, q, mS)
    g = torch.autograd.grad(E, [p, b, q])
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    assert rel(g[0], go[0]) < RTOL and rel(g[2], go[2]) < RTOL
    assert rel(torch.diagonal(g[1]), torch.diagonal(go[1])) < RTOL


# ------------------------------------------------------------------------------------------ multipolar PME
def test_nonpol_energy_and_all_gradients(carved):
    from admp_b200.pme import ADMPPmeForce
    s, pairs = carved
    Ql, _, _, _ = _perturbed(s)
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2)
    assert (calc.kappa, calc.K1, calc.K2, calc.K3) == (ref.kappa, ref.K1, ref.K2, ref.K3)
    mS0 = [0.1, 0.3, 0.0, 0.7, 1.0]
    tp, tb, tQ, tm = _t(s.positions), _t(s.box), _t(Ql), _t(mS0)
    parts = {}
    Eo = ref.get_energy(tp, tb, pairs, tQ, tm, parts=parts)
    go = torch.autograd.grad(Eo, [tp, tb, tQ, tm])
    p, b, q, m = (torch.tensor(np.asarray(x), device='cuda', dtype=torch.float64, requires_grad=True)
                  for x in (s.positions, s.box, Ql, mS0))
    E = calc.get_energy(p, b, pairs, q, m)
    g = torch.autograd.grad(E, [p, b, q, m])
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item()), (E.item(), Eo.item(), parts)
    assert rel(g[0], go[0]) < RTOL
    assert rel(g[2], go[2]) < RTOL
    assert rel(g[3], go[3]) < RTOL
    assert rel(g[1], go[1]) < 1e-5      # virial: off-diagonals carry Nyquist-mode conventions
    assert rel(torch.diagonal(g[1]), torch.diagonal(go[1])) < RTOL
    # get_forces == (E, +dE/dpositions)  
    E2, F2 = calc.get_forces(s.positions, s.box, pairs, Ql, mS0)
    assert abs(E2.item() - Eo.item()) < RTOL * abs(Eo.item()) and rel(F2, go[0]) < RTOL


def test_polarizable_energy_fn_and_all_gradients(carved):
    from admp_b200.pme import ADMPPmeForce
    s, pairs = carved
    Ql, U, pol, th = _perturbed(s)
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2, lpol=True)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2, lpol=True)
    mS0, pS0, dS0 = [0.1, 0.3, 0.0, 0.7, 1.0], [0.0, 0.4, 0.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0, 1.0]
    names = ['pos', 'box', 'Ql', 'U', 'pol', 'th', 'mS', 'pS']
    vals = [s.positions, s.box, Ql, U, pol, th, mS0, pS0]
    to = [_t(v) for v in vals]
    Eo = ref.energy_fn(to[0], to[1], pairs, to[2], to[3], to[4], to[5], to[6], to[7], _t(dS0, False))
    go = torch.autograd.grad(Eo, to)
    tg = [torch.tensor(np.asarray(v), device='cuda', dtype=torch.float64, requires_grad=True) for v in vals]
    E = calc.energy_fn(tg[0], tg[1], pairs, tg[2], tg[3], tg[4], tg[5], tg[6], tg[7], dS0)
    g = torch.autograd.grad(E, tg)
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    for k, nm in enumerate(names):
        if nm == 'box':
            assert rel(torch.diagonal(g[k]), torch.diagonal(go[k])) < RTOL, nm
            assert rel(g[k], go[k]) < 1e-5, nm
        else:
            assert rel(g[k], go[k]) < RTOL, (nm, rel(g[k], go[k]))
    # grad_U_fn / grad_pos_fn closures
    assert rel(calc.grad_U_fn(*[vals[0], vals[1], pairs, vals[2], vals[3], vals[4], vals[5], mS0, pS0, dS0]), go[3]) < RTOL
    assert rel(calc.grad_pos_fn(*[vals[0], vals[1], pairs, vals[2], vals[3], vals[4], vals[5], mS0, pS0, dS0]), go[0]) < RTOL


@pytest.mark.parametrize('lmax', [0, 1])
def test_lower_multipole_orders_nonpol(carved, lmax):
    """lmax = 0 (charges only; admp/pme.py:224-231 skips the frames) and lmax = 1 (charges + dipoles): Q_local has
    (lmax+1)^2 columns, energy and every gradient against the oracle."""
    from admp_b200.pme import ADMPPmeForce
    s, pairs = carved
    nh = (lmax + 1) ** 2
    Ql = _perturbed(s)[0][:, :nh].copy()
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, lmax)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, lmax)
    mS0 = [0.1, 0.3, 0.0, 0.7, 1.0]
    tp, tb, tQ, tm = _t(s.positions), _t(s.box), _t(Ql), _t(mS0)
    Eo = ref.get_energy(tp, tb, pairs, tQ, tm)
    go = torch.autograd.grad(Eo, [tp, tb, tQ, tm])
    p, b, q, m = (torch.tensor(np.asarray(x), device='cuda', dtype=torch.float64, requires_grad=True)
                  for x in (s.positions, s.box, Ql, mS0))
    E = calc.get_energy(p, b, pairs, q, m)
    g = torch.autograd.grad(E, [p, b, q, m])
    assert g[2].shape == (s.n_atoms, nh)
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item()), (E.item(), Eo.item())
    assert rel(g[0], go[0]) < RTOL and rel(g[2], go[2]) < RTOL and rel(g[3], go[3]) < RTOL
    assert rel(torch.diagonal(g[1]), torch.diagonal(go[1])) < RTOL


def test_dipole_order_polarizable(carved):
    """lmax = 1 with induced dipoles: energy_fn and its gradients (positions, Q_local, U, pol, tholes, scales)."""
    from admp_b200.pme import ADMPPmeForce
    s, pairs = carved
    Ql, U, pol, th = _perturbed(s)
    Ql = Ql[:, :4].copy()
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 1, lpol=True)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 1, lpol=True)
    mS0, pS0, dS0 = [0.1, 0.3, 0.0, 0.7, 1.0], [0.0, 0.4, 0.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0, 1.0]
    vals = [s.positions, s.box, Ql, U, pol, th, mS0, pS0]
    to = [_t(v) for v in vals]
    Eo = ref.energy_fn(to[0], to[1], pairs, to[2], to[3], to[4], to[5], to[6], to[7], _t(dS0, False))
    go = torch.autograd.grad(Eo, to)
    tg = [torch.tensor(np.asarray(v), device='cuda', dtype=torch.float64, requires_grad=True) for v in vals]
    E = calc.energy_fn(tg[0], tg[1], pairs, tg[2], tg[3], tg[4], tg[5], tg[6], tg[7], dS0)
    g = torch.autograd.grad(E, tg)
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    for k in (0, 2, 3, 4, 5, 6, 7):
        assert rel(g[k], go[k]) < RTOL, (k, rel(g[k], go[k]))
    assert rel(torch.diagonal(g[1]), torch.diagonal(go[1])) < RTOL


def test_scf_converging_system_matches_oracle_iteration_for_iteration(lattice):
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    for thresh in (10.0, 1e-3):
        Uo, fo, no = ref.optimize_Uind(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales,
                                       s.dScales, thresh=thresh)
        U, f, n = calc.optimize_Uind(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales,
                                     s.dScales, thresh=thresh)
        assert (f, n) == (fo, no), (thresh, f, n, fo, no)
        assert rel(U, Uo) < RTOL
    assert calc._ctx.scf_graph_active, 'device-resident SCF graph was not used'
    # full get_forces (SCF from U=0, energy and gradient at fixed U) and warm start
    tp = _t(s.positions)
    Eo = ref.get_energy(tp, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                        U_init=torch.zeros(s.n_atoms, 3, dtype=torch.float64))
    go = torch.autograd.grad(Eo, tp)[0]
    E, F = calc.get_forces(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    assert calc.n_cycle == ref.n_cycle and calc.lconverg == ref.lconverg
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    assert rel(F, go) < RTOL and rel(calc.U_ind, ref.U_ind) < RTOL
    E2, _ = calc.get_forces(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                            U_init=calc.U_ind)
    Eo2 = ref.get_energy(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                         U_init=ref.U_ind)
    assert calc.n_cycle == ref.n_cycle == 0
    assert abs(E2.item() - Eo2.item()) < RTOL * abs(Eo2.item())


def test_scf_graph_loop_equals_host_synchronised_loop_and_handles_non_convergence(carved):
    """On the gas-like carved box the reference's Jacobi iteration DIVERGES (O-O contacts of ~1.1 A):
    30 iterations, flag False - reproduced iteration for iteration."""
    from admp_b200.pme import ADMPPmeForce
    from admp_b200 import _lib
    s, pairs = carved
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2, lpol=True)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 6.0, 1e-4, 2, lpol=True)
    Uo, fo, no = ref.optimize_Uind(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                                   maxiter=8)
    U, f, n = calc.optimize_Uind(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                                 maxiter=8)
    assert (f, n) == (fo, no)
    assert rel(U, Uo) < RTOL
    args = [calc._prep(x) for x in (s.positions, s.box, s.Q_local)]
    rest = [calc._prep(x) for x in (s.pol, s.tholes, s.mScales, s.pScales)]
    pr = torch.as_tensor(pairs, device='cuda')
    a = calc._eval(args[0], args[1], pr, args[2], None, *rest, _lib.WANT_GRAD, True, maxiter=8, cache_scf=False)
    b = calc._eval(args[0], args[1], pr, args[2], None, *rest, _lib.WANT_GRAD, True, maxiter=8, hostsync=True, cache_scf=False)
    assert torch.equal(a.scf, b.scf)
    assert rel(a.U, b.U) < 1e-12 and rel(a.dpos, b.dpos) < 1e-10
    assert abs(a.energy.item() - b.energy.item()) < 1e-10 * abs(b.energy.item())


@pytest.mark.parametrize('sides, rc', [((8, 8, 8), 8.0), ((6, 7, 12), 7.5), ((3, 3, 20), 4.6)])
def test_neighbor_list_liquid_density_many_neighbours_per_atom(sides, rc):
    """Warp-per-atom list kernels on boxes where an atom has ~100 listed partners (several 32-candidate rounds, rank sort of a long
    row segment) and where a dimension holds exactly 3 or only 2 cells: pair array identical to the oracle's, row for row."""
    from admp_b200 import workloads
    from admp_b200.neighbor import neighbor_list
    w = workloads.dense_water(sides)
    po, no = pairlist.build_pairs(np.asarray(w.positions), np.asarray(w.box), rc)
    nbr = neighbor_list(w.box, rc).allocate(w.positions)
    assert nbr.n_pairs == no and not nbr.did_buffer_overflow
    assert np.array_equal(nbr.pairs[:no].cpu().numpy(), po[:no])
    assert 2.0 * no / w.n_atoms > 30


def test_reference_example_water_1024_nonpol():
    """Config C1: examples/water_1024 (3072 atoms, rc 4, kappa 0.657065221219616, K 154^3)."""
    from admp_b200.pme import ADMPPmeForce
    from admp_b200.neighbor import neighbor_list
    s = fixtures.water1024().nonpol()
    pairs_o, n_o = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.0)
    nbr = neighbor_list(s.box, 4.0).allocate(s.positions)
    assert nbr.n_pairs == n_o == 12272 and not nbr.did_buffer_overflow
    assert np.array_equal(nbr.pairs[:n_o].cpu().numpy(), pairs_o[:n_o]), 'pair set must be bit-exact'
    assert bool((nbr.pairs[n_o:] == s.n_atoms).all())
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2)
    calc.update_env('kappa', fixtures.KAPPA_EXAMPLE)
    assert (calc.K1, calc.K2, calc.K3) == (154, 154, 154)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2)
    ref.update_env('kappa', fixtures.KAPPA_EXAMPLE)
    tp = _t(s.positions)
    Eo = ref.get_energy(tp, s.box, pairs_o, s.Q_local, s.mScales)
    go = torch.autograd.grad(Eo, tp)[0]
    E, F = calc.get_forces(s.positions, s.box, nbr.pairs, s.Q_local, s.mScales)
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item()), (E.item(), Eo.item())
    assert rel(F, go) < RTOL


def test_neighbor_list_skin_update_semantics():
    """jax_md's allocate / update contract: a list built with a skin is reused while no atom
    moved further than dr_threshold / 2 and still contains every pair inside rc; a larger move rebuilds it and the
    rebuilt pair set is again bit-exact against the oracle's predicate."""
    from admp_b200.neighbor import neighbor_list
    s = fixtures.water1024().nonpol()
    rc, skin = 4.0, 0.6
    fn = neighbor_list(s.box, rc, dr_threshold=skin)
    nbr = fn.allocate(s.positions)
    pairs_skin, n_skin = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), rc + skin)
    assert nbr.n_pairs == n_skin and np.array_equal(nbr.pairs[:n_skin].cpu().numpy(), pairs_skin[:n_skin])
    rng = np.random.default_rng(5)
    small = s.positions.numpy() + rng.uniform(-0.1, 0.1, size=s.positions.shape)        # |dr| <= 0.173 < skin / 2
    same = nbr.update(small)
    assert same is nbr
    exact, n_exact = pairlist.build_pairs(small, s.box.numpy(), rc)
    listed = set(map(tuple, nbr.pairs[:n_skin].cpu().numpy().tolist()))
    assert all(tuple(p) in listed for p in exact[:n_exact].tolist()), 'the skinned list lost a pair inside rc'
    big = small.copy()
    big[7] += np.array([0.0, 0.5, 0.0])                                                  # one atom beyond skin / 2
    rebuilt = nbr.update(big)
    assert rebuilt is not nbr and not rebuilt.did_buffer_overflow
    ref, n_ref = pairlist.build_pairs(big, s.box.numpy(), rc + skin)
    assert rebuilt.n_pairs == n_ref and np.array_equal(rebuilt.pairs[:n_ref].cpu().numpy(), ref[:n_ref])


def test_replica_invariance_polarizable(lattice):
    """with K scaled by the replication factors, E(replicated) = n_rep * E and forces and
    induced dipoles replicate - the parity check at sizes the reference could never run."""
    from admp_b200.pme import ADMPPmeForce
    from admp_b200.neighbor import neighbor_list
    s, _ = lattice
    out = []
    for rep in ((1, 1, 1), (2, 1, 3)):
        r = s.replicate(*rep)
        calc = ADMPPmeForce(s.box, r.axis_type, r.axis_indices, r.covalent_map, 5.0, 1e-4, 2, lpol=True)
        K0 = (calc.K1, calc.K2, calc.K3)
        for d in range(3):
            calc.update_env('K%d' % (d + 1), K0[d] * rep[d])
        pairs = neighbor_list(r.box, 5.0).allocate(r.positions).pairs
        E, F = calc.get_forces(r.positions, r.box, pairs, r.Q_local, r.pol, r.tholes, r.mScales, r.pScales, r.dScales)
        out.append((E.item(), F.cpu(), calc.U_ind.cpu(), calc.n_cycle))
    nrep = 6
    assert out[0][3] == out[1][3]
    assert abs(out[1][0] - nrep * out[0][0]) < 1e-8 * abs(nrep * out[0][0])
    assert rel(out[1][1], out[0][1].repeat(nrep, 1)) < 1e-7
    assert rel(out[1][2], out[0][2].repeat(nrep, 1)) < 1e-7


def test_generate_pme_recip_matches_oracle(carved):
    from admp_b200.recip import generate_pme_recip, Ck_1, Ck_6
    s, _ = carved
    Ql, _, _, _ = _perturbed(s)
    fr = construct_local_frames(s.positions, s.box, s.axis_type, s.axis_indices)
    Qg = rot_local2global(torch.tensor(Ql), fr, 2)
    kappa, K = 0.45, (40, 36, 50)
    fn = generate_pme_recip(Ck_1, kappa, False, 6, K[0], K[1], K[2], 2)
    tp, tb, tq = _t(s.positions), _t(s.box), _t(Qg)
    Eo = orecip.pme_recip(tp, tb, tq, kappa, K, 2, kind=1, gamma=False)
    go = torch.autograd.grad(Eo, [tp, tb, tq])
    p, b, q = (torch.tensor(np.asarray(x.detach()), device='cuda', requires_grad=True) for x in (tp, tb, tq))
    E = fn(p, b, q)
    g = torch.autograd.grad(E, [p, b, q])
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    assert rel(g[0], go[0]) < RTOL and rel(g[2], go[2]) < RTOL and rel(g[1], go[1]) < 1e-5
    fn6 = generate_pme_recip(Ck_6, kappa, True, 6, K[0], K[1], K[2], 0)
    c6 = s.c_list[:, 0:1]
    Eo6 = orecip.pme_recip(s.positions, s.box, c6, kappa, K, 0, kind=6, gamma=True)
    assert abs(fn6(s.positions, s.box, c6).item() - Eo6.item()) < RTOL * abs(Eo6.item())


# ------------------------------------------------------------------------------------------ dispersion / TT
@pytest.mark.parametrize('pmax', [6, 8, 10])
def test_dispersion_pme(carved, pmax):
    from admp_b200.disp_pme import ADMPDispPmeForce
    s, pairs = carved
    calc = ADMPDispPmeForce(s.box, s.covalent_map, 6.0, 1e-4, pmax)
    ref = OracleDispPmeForce(s.box, s.covalent_map, 6.0, 1e-4, pmax)
    mS0 = [0.1, 0.3, 0.0, 0.7, 1.0]
    tp, tb, tc, tm = _t(s.positions), _t(s.box), _t(s.c_list), _t(mS0)
    Eo = ref.get_energy(tp, tb, pairs, tc, tm)
    go = torch.autograd.grad(Eo, [tp, tc, tm])
    p, b, c, m = (torch.tensor(np.asarray(x), device='cuda', dtype=torch.float64, requires_grad=True)
                  for x in (s.positions, s.box, s.c_list, mS0))
    E = calc.get_energy(p, b, pairs, c, m)
    g = torch.autograd.grad(E, [p, c, m])
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    assert rel(g[0], go[0]) < RTOL
    ncol = (pmax - 4) // 2
    assert rel(g[1][:, :ncol], go[1][:, :ncol]) < RTOL
    assert rel(g[2], go[2]) < RTOL
    E2, F2 = calc.get_forces(s.positions, s.box, pairs, s.c_list, mS0)
    assert rel(F2, go[0]) < RTOL


def test_tt_pair_interaction(carved):
    from admp_b200.pairwise import generate_pairwise_interaction, TT_damping_qq_c6_kernel
    s, pairs = carved
    fn = generate_pairwise_interaction(TT_damping_qq_c6_kernel, s.covalent_map, static_args={})
    ofn = o_pairwise(o_tt, s.covalent_map, {})
    mS0 = [0.1, 0.3, 0.0, 0.7, 1.0]
    vals = [s.positions, s.box, mS0, s.tt_a, s.tt_b, s.tt_q, s.c_list[:, 0]]
    to = [_t(v) for v in vals]
    Eo = ofn(to[0], to[1], pairs, to[2], *to[3:])
    go = torch.autograd.grad(Eo, [to[0]] + to[2:])
    tg = [torch.tensor(np.asarray(v), device='cuda', dtype=torch.float64, requires_grad=True) for v in vals]
    E = fn(tg[0], tg[1], pairs, tg[2], *tg[3:])
    g = torch.autograd.grad(E, [tg[0]] + tg[2:])
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    for a, b in zip(g, go):
        assert rel(a, b) < RTOL


# ------------------------------------------------------------------------------------------ single precision
def test_single_precision_within_1e4(lattice):
    from admp_b200 import settings
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    tp = _t(s.positions)
    Eo = ref.get_energy(tp, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales,
                        U_init=torch.zeros(s.n_atoms, 3, dtype=torch.float64))
    go = torch.autograd.grad(Eo, tp)[0]
    old = settings.PRECISION
    settings.PRECISION = 'single'
    try:
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
        E, F = calc.get_forces(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        assert E.dtype == torch.float32 and F.dtype == torch.float32
        assert calc.n_cycle == ref.n_cycle
        assert abs(E.item() - Eo.item()) < 1e-4 * abs(Eo.item())
        assert rel(F, go) < 1e-4
    finally:
        settings.PRECISION = old


# ------------------------------------------------------------------------------------------ edge cases
def test_empty_and_padded_pair_lists(lattice):
    """Only rows with pairs[:,0] < pairs[:,1] are used (admp/pme.py:671): (N,N) padding, reversed and
    self rows vanish; an empty list leaves reciprocal + self."""
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2)
    n = s.n_atoms
    junk = np.array([[n, n], [5, 2], [7, 7], [n, n]], dtype=np.int32)
    padded = np.concatenate([pairs[:10], junk, pairs[10:], junk])
    E0 = calc.get_energy(s.positions, s.box, pairs, s.Q_local, s.mScales).item()
    E1 = calc.get_energy(s.positions, s.box, padded, s.Q_local, s.mScales).item()
    assert E0 == pytest.approx(E1, rel=1e-10)      # atomic accumulation order differs
    empty = np.zeros((0, 2), dtype=np.int32)
    Ee = calc.get_energy(s.positions, s.box, empty, s.Q_local, s.mScales).item()
    Eo = ref.get_energy(s.positions, s.box, np.full((1, 2), n), s.Q_local, s.mScales).item()
    assert abs(Ee - Eo) < RTOL * abs(Eo)


def test_neighbor_list_capacity_overflow_and_skin(lattice):
    from admp_b200.neighbor import neighbor_list
    s, _ = lattice
    fn = neighbor_list(s.box, 5.0)
    nbr = fn.allocate(s.positions)
    po, no = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 5.0)
    assert nbr.n_pairs == no and np.array_equal(nbr.pairs[:no].cpu().numpy(), po[:no])
    assert nbr.idx.shape[0] == 2
    small = fn._build(s.positions, no // 2)
    assert small.did_buffer_overflow and small.n_pairs == no
    assert np.array_equal(small.pairs.cpu().numpy(), po[:no // 2])
    moved = s.positions + 0.3
    pm, nm = pairlist.build_pairs(moved.numpy(), s.box.numpy(), 5.0)
    upd = nbr.update(moved)
    assert upd.n_pairs == nm and np.array_equal(upd.pairs[:nm].cpu().numpy(), pm[:nm])


# ------------------------------------------------------------------------------------------ module-level surface
def test_module_level_functions_of_the_reference_surface(carved):
    """admp/pme.py and admp/disp_pme.py expose their building blocks as free functions (energy_pme, pme_real,
    pme_self, pol_penalty, energy_disp_pme, disp_pme_real, disp_pme_self, g_p): same names, argument order and
    values here, against the oracle's restatements."""
    from admp_b200 import pme as P, disp_pme as D
    from admp_b200.spatial import generate_construct_local_frames
    from oracle.frames import construct_local_frames as o_frames
    from oracle.harmonics import rot_local2global as o_l2g, cart_dipole_to_harm
    from oracle import dispersion as odisp
    s, pairs = carved
    Ql, U, pol, th = _perturbed(s)
    mS0, pS0, dS0 = [0.1, 0.3, 0.0, 0.7, 1.0], [0.0, 0.4, 0.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0, 1.0]
    kappa, K = 0.41, 30
    # energy_pme, polarizable and not, with its position gradient
    fn = generate_construct_local_frames(s.axis_type, s.axis_indices)
    to = [_t(v) for v in (s.positions, s.box, Ql, U, pol, th, mS0, pS0)]
    Eo = orc.energy_pme(to[0], to[1], pairs, to[2], to[3], to[4], to[5], to[6], to[7], _t(dS0, False), s.covalent_map,
                        s.axis_type, s.axis_indices, kappa, K, K, K, 2, True)
    go = torch.autograd.grad(Eo, to[0])[0]
    pos = torch.tensor(s.positions.numpy(), device='cuda', requires_grad=True)
    E = P.energy_pme(pos, s.box, pairs, Ql, U, pol, th, mS0, pS0, dS0, s.covalent_map, fn, None, kappa, K, K, K, 2, True)
    assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
    assert rel(torch.autograd.grad(E, pos)[0], go) < RTOL
    En = P.energy_pme(s.positions, s.box, pairs, Ql, None, None, None, mS0, None, None, s.covalent_map, fn, None, kappa, K, K, K, 2, False)
    Eno = orc.energy_pme(to[0], to[1], pairs, to[2], None, None, None, to[6], None, None, s.covalent_map,
                         s.axis_type, s.axis_indices, kappa, K, K, K, 2, False)
    assert abs(En.item() - Eno.item()) < RTOL * abs(Eno.item())
    # pme_real from global-frame harmonic multipoles and harmonic-order induced dipoles
    Qg = o_l2g(_t(Ql, False), o_frames(s.positions, s.box, s.axis_type, s.axis_indices), 2)
    Uh = cart_dipole_to_harm(_t(U, False))
    qg, uh, tp = Qg.clone().requires_grad_(True), Uh.clone().requires_grad_(True), _t(s.positions)
    Ero = orc.pme_real(tp, _t(s.box, False), pairs, qg, uh, _t(pol, False), _t(th, False), _t(mS0, False), _t(pS0, False), _t(dS0, False),
                       s.covalent_map, kappa, 2, True)
    gro = torch.autograd.grad(Ero, [tp, qg, uh])
    a = [torch.tensor(x.detach().numpy(), device='cuda', requires_grad=True) for x in (tp, qg, uh)]
    Er = P.pme_real(a[0], s.box, pairs, a[1], a[2], pol, th, mS0, pS0, dS0, s.covalent_map, kappa, 2, True)
    gr = torch.autograd.grad(Er, a)
    assert abs(Er.item() - Ero.item()) < RTOL * abs(Ero.item())
    for k in range(3):
        assert rel(gr[k], gro[k]) < RTOL, k
    # closed forms
    assert abs(P.pme_self(Qg, kappa, 2).item() - orc.pme_self(Qg, kappa, 2).item()) < 1e-12 * abs(orc.pme_self(Qg, kappa, 2).item())
    assert abs(P.pol_penalty(Uh, _t(pol, False)).item() - orc.pol_penalty(Uh, _t(pol, False)).item()) < 1e-9 * orc.pol_penalty(Uh, _t(pol, False)).item()
    # dispersion
    for pmax in (6, 10):
        Ed = D.energy_disp_pme(s.positions, s.box, pairs, s.c_list, mS0, s.covalent_map, kappa, K, K, K, pmax)
        Edo = odisp.energy_disp_pme(s.positions, s.box, pairs, s.c_list, _t(mS0, False), s.covalent_map, kappa, K, K, K, pmax)
        assert abs(Ed.item() - Edo.item()) < RTOL * abs(Edo.item())
    x2 = torch.tensor([0.3, 1.7, 4.0], dtype=torch.float64)
    g = D.g_p(x2, 10)
    ref = torch.exp(-x2) * torch.stack([1 + x2 + x2 ** 2 / 2, 1 + x2 + x2 ** 2 / 2 + x2 ** 3 / 6,
                                        1 + x2 + x2 ** 2 / 2 + x2 ** 3 / 6 + x2 ** 4 / 24])
    assert torch.allclose(g, ref, rtol=1e-14)
    Eself = D.disp_pme_self(s.c_list, kappa, 10).item()
    c = s.c_list
    assert abs(Eself - (-kappa ** 6 / 12 * (c[:, 0] ** 2).sum() - kappa ** 8 / 48 * (c[:, 1] ** 2).sum()
                        - kappa ** 10 / 240 * (c[:, 2] ** 2).sum()).item()) < 1e-12 * abs(Eself)
    Edr = D.disp_pme_real(s.positions, s.box, pairs, s.c_list, mS0, s.covalent_map, kappa, 10)
    Edro = odisp.disp_pme_real(s.positions, s.box, pairs, s.c_list, _t(mS0, False), s.covalent_map, kappa, 10)
    assert abs(Edr.item() - Edro.item()) < RTOL * abs(Edro.item())
