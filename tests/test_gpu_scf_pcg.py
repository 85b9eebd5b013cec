"""GPU tests of the beyond-reference SCF solver (settings.SCF_SOLVER = 'pcg', ADMP_SCF_CG): the CUDA
conjugate-gradient loop (site.cu scf_cg_kernel inside the same CUDA-graph WHILE body as the Jacobi loop) against its CPU
restatement (oracle/scf_pcg.py) - iteration counts identical, dipoles within 1e-6 relative - and against the fixed point of
the reference's Jacobi iteration.  The default solver stays the reference's."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fixtures, pairlist                                   # noqa: E402
from oracle import realspace as orc                                     # noqa: E402
from oracle.scf_pcg import optimize_Uind_pcg                            # noqa: E402

RTOL = 1e-6


def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


@pytest.fixture(scope='module')
def lattice():
    s = fixtures.lattice_water(4, 3.15, seed=5)  # 64 waters, 12.6 A box, liquid-like (SCF converges)
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 5.0)
    return s, pairs


def test_pcg_matches_its_oracle_iteration_for_iteration_and_the_jacobi_fixed_point(lattice):
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    args = (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    for thresh in (10.0, 1e-3, 1e-7):
        Uo, fo, no, _ = optimize_Uind_pcg(ref, *args, maxiter=60, thresh=thresh)
        U, f, n = calc.optimize_Uind(*args, maxiter=60, thresh=thresh, solver='pcg')
        assert (f, n) == (fo, no), (thresh, f, n, fo, no)
        assert rel(U, Uo) < RTOL
    assert calc._ctx.scf_graph_active, 'device-resident SCF graph was not used'
    # same fixed point as the reference's iteration, in fewer passes
    Uj, fj, nj = calc.optimize_Uind(*args, maxiter=60, thresh=1e-7)
    assert fj and f and n + 2 < nj + 1
    assert rel(U, Uj) < 1e-7
    # warm start on the converged dipoles: zero iterations
    U2, f2, n2 = calc.optimize_Uind(*args, U_init=U, maxiter=60, thresh=1e-6, solver='pcg')
    assert f2 and n2 == 0 and rel(U2, U) < 1e-14
    # iteration budget: stops after maxiter CG iterations, flag False (true residual still above the threshold)
    U3, f3, n3 = calc.optimize_Uind(*args, maxiter=2, thresh=1e-9, solver='pcg')
    Uo3, fo3, no3, _ = optimize_Uind_pcg(ref, *args, maxiter=2, thresh=1e-9)
    assert (f3, n3) == (fo3, no3) == (False, 2) and rel(U3, Uo3) < RTOL


def test_pcg_through_the_public_energy_and_force_calls(lattice):
    """settings.SCF_SOLVER = 'pcg': get_forces = SCF by CG, then energy and gradient at fixed U (Hellmann-Feynman), against
    the oracle's energy function evaluated on the oracle solver's dipoles; graph loop == host-synchronised loop."""
    from admp_b200 import settings, _lib
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    args = (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    old = settings.SCF_SOLVER, settings.POL_CONV
    try:
        settings.SCF_SOLVER, settings.POL_CONV = 'pcg', 1e-6
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
        E, g = calc.get_forces(*args)
        Uo, fo, no, _ = optimize_Uind_pcg(ref, *args, thresh=1e-6)
        pos = s.positions.clone().requires_grad_(True)
        Eo = ref.energy_fn(pos, s.box, pairs, s.Q_local, Uo, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        go = torch.autograd.grad(Eo, pos)[0]
        assert (calc.lconverg, calc.n_cycle) == (fo, no)
        assert abs(E.item() - Eo.item()) < RTOL * abs(Eo.item())
        assert rel(g, go) < RTOL and rel(calc.U_ind, Uo) < RTOL
        # with the virial (dE/dbox) the final reciprocal pass runs after the loop: same energy
        E2, g2, vir = calc.get_forces_and_virial(*args)
        assert abs(E2.item() - E.item()) < 1e-10 * abs(E.item()) and rel(g2, g) < 1e-9
        a4 = [calc._prep(x) for x in (s.positions, s.box, s.Q_local)]
        rest = [calc._prep(x) for x in (s.pol, s.tholes, s.mScales, s.pScales)]
        pr = torch.as_tensor(pairs, device='cuda')
        a = calc._eval(a4[0], a4[1], pr, a4[2], None, *rest, _lib.WANT_GRAD, True, cache_scf=False)
        b = calc._eval(a4[0], a4[1], pr, a4[2], None, *rest, _lib.WANT_GRAD, True, hostsync=True, cache_scf=False)
        assert torch.equal(a.scf, b.scf)
        assert rel(a.U, b.U) < 1e-12 and rel(a.dpos, b.dpos) < 1e-10
    finally:
        settings.SCF_SOLVER, settings.POL_CONV = old
    # back on the default: the reference's iteration, cycle for cycle
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    U, f, n = calc.optimize_Uind(*args)
    Uj, fj, nj = ref.optimize_Uind(*args)
    assert (f, n) == (fj, nj) and rel(U, Uj) < RTOL


def test_pcg_on_the_headline_box_reports_the_indefinite_matrix():
    """Config C2 (the shipped 1024-water box): the Jacobi loop runs its 30 cycles and diverges; CG finds a direction of
    negative curvature after a few iterations and stops (flag False), as its oracle does - same iteration, same dipoles."""
    from admp_b200.pme import ADMPPmeForce
    s = fixtures.water1024()
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.0)
    ref = orc.OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2, lpol=True)
    ref.update_env('kappa', fixtures.KAPPA_EXAMPLE)
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2, lpol=True)
    calc.update_env('kappa', fixtures.KAPPA_EXAMPLE)
    args = (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    Uo, fo, no, _ = optimize_Uind_pcg(ref, *args)
    U, f, n = calc.optimize_Uind(*args, solver='pcg')
    assert (f, n) == (fo, no) and not f and n < 12
    assert rel(U, Uo) < 1e-5


def test_unknown_solver_is_rejected(lattice):
    from admp_b200.pme import ADMPPmeForce
    s, pairs = lattice
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    with pytest.raises(ValueError):
        calc.optimize_Uind(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales, solver='diis')
