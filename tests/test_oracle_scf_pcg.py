"""CPU checks of the beyond-reference SCF solver's restatement (oracle/scf_pcg.py): it reaches the
fixed point of the reference's Jacobi iteration (admp/pme.py:132-138) in fewer field evaluations, solves the dense linear
system, and reports an indefinite polarization matrix instead of diverging."""
import numpy as np
import torch

from oracle import fixtures, pairlist
from oracle.realspace import OraclePmeForce, DIELECTRIC
from oracle.scf_pcg import optimize_Uind_pcg


def _lattice():
    s = fixtures.lattice_water(3, 3.2, seed=11)
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.0)
    f = OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2, lpol=True)
    return s, f, (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)


def test_pcg_reaches_the_jacobi_fixed_point_in_fewer_field_evaluations():
    s, f, args = _lattice()
    Uj, fj, nj = f.optimize_Uind(*args, maxiter=60, thresh=1e-7)
    U, flag, it, n_field = optimize_Uind_pcg(f, *args, maxiter=60, thresh=1e-7)
    assert fj and flag
    assert (U - Uj).abs().max().item() < 1e-9 * Uj.abs().max().item() + 1e-10
    assert n_field == it + 2 and n_field < nj + 1          # Jacobi: nj + 1 field evaluations
    # the stopping rule is the reference's, evaluated on the returned U
    F = f.grad_U_fn(*args[:4], U, *args[4:])
    assert F[s.pol > 0.001].abs().max().item() < 1e-7
    # loose default threshold: both stop early, on different iterates of the same sequence family
    U10, flag10, it10, _ = optimize_Uind_pcg(f, *args)
    assert flag10 and it10 <= 3


def test_pcg_solves_the_dense_linear_system():
    s, f, args = _lattice()
    n = s.n_atoms
    zero = torch.zeros(n, 3, dtype=torch.float64)
    field = lambda u: f.grad_U_fn(*args[:4], u, *args[4:])
    b = -field(zero)
    cols = []
    for k in range(3 * n):                                  # A e_k = field(e_k) - field(0)
        e = torch.zeros(3 * n, dtype=torch.float64)
        e[k] = 1.0
        cols.append((field(e.reshape(n, 3)) + b).reshape(-1))
    A = torch.stack(cols, 1)
    assert (A - A.T).abs().max().item() < 1e-8 * A.abs().max().item()
    pol3 = s.pol.repeat_interleave(3)
    keep = pol3 > 0
    Ud = torch.zeros(3 * n, dtype=torch.float64)
    Ud[keep] = torch.linalg.solve(A[keep][:, keep], b.reshape(-1)[keep])
    U, flag, it, _ = optimize_Uind_pcg(f, *args, maxiter=80, thresh=1e-9)
    assert flag
    assert (U.reshape(-1) - Ud).abs().max().item() < 1e-8 * Ud.abs().max().item()
    # Jacobi converges here because every eigenvalue of the iteration matrix lies inside (-1, 1)
    Mh = torch.sqrt(pol3[keep] / DIELECTRIC)
    lam = torch.linalg.eigvalsh(torch.eye(int(keep.sum()), dtype=torch.float64) - Mh[:, None] * A[keep][:, keep] * Mh[None, :])
    assert lam.abs().max().item() < 1.0


def test_pcg_reports_an_indefinite_matrix_instead_of_diverging():
    """The shipped 1024-water box (config C2) has O-O contacts of ~1.1 A: the Jacobi iteration diverges on it
    because the polarization matrix is not positive definite; CG meets a direction of negative curvature after a few
    iterations, stops there and says so (flag False, bounded dipoles)."""
    s = fixtures.water1024()
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.0)
    f = OraclePmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 4.0, 1e-4, 2, lpol=True)
    f.update_env('kappa', fixtures.KAPPA_EXAMPLE)
    args = (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    U, flag, it, n_field = optimize_Uind_pcg(f, *args, maxiter=30)
    Uj, fj, nj = f.optimize_Uind(*args, maxiter=12)
    assert not flag and not fj
    assert it < 12 and n_field == it + 3 and np.isfinite(U.numpy()).all()
    assert U.abs().max().item() < 10.0
    # the Jacobi iterate is further from a fixed point after 12 cycles than after 6: the residual grows
    U6, _, _ = f.optimize_Uind(*args, maxiter=6)
    site = s.pol > 0.001
    res = lambda u: f.grad_U_fn(*args[:4], u, *args[4:])[site].abs().max().item()
    assert res(Uj) > res(U6)
