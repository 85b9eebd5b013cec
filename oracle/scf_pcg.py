"""TEST INFRASTRUCTURE (CPU oracle) - beyond-reference SCF solver, SURVEY 8(f) rank 4 ("better SCF: preconditioned CG").

The reference converges the induced dipoles with a Jacobi iteration (admp/pme.py:111-143:
U <- U - field * pol / DIELECTRIC, test max|field| over pol > 0.001 BEFORE the update).  The field is
affine in U, field(U) = A U - b with A symmetric, so the same fixed point is the solution of a linear system
and conjugate gradients preconditioned with the Jacobi scaling M = pol / DIELECTRIC reach it in far fewer
field evaluations - and also when the Jacobi iteration matrix has eigenvalues below -1 (where Jacobi diverges
although A is positive definite).

This file states the algorithm of the CUDA path (admp_b200/csrc/site.cu scf_cg_kernel) step for step on top of
the oracle's field function; tests/ compare the two.  Only the field function is shared with the reference
(OraclePmeForce.grad_U_fn = jax.grad(energy_fn, argnums=4), admp/pme.py:76); the solver itself has no reference
counterpart ("parity unpinned" by construction: it is checked against the Jacobi fixed point and a dense solve).

Work per call: one field evaluation for the starting residual, one per CG iteration (at the trial point U + p:
A p = field(U + p) - field(U)), one at the end on the final U (the TRUE residual decides, and the reciprocal
mesh of the final U is what the energy / force pass needs).  If the true residual fails the test the iteration
restarts from it.
"""
import torch

from .realspace import DIELECTRIC, MAX_N_POL, POL_CONV


def optimize_Uind_pcg(force, positions, box, pairs, Q_local, pol, tholes, mScales, pScales, dScales,
                      U_init=None, maxiter=MAX_N_POL, thresh=POL_CONV):
    """Returns (U, flag, n_iter, n_field): flag True when max|field(U)| < thresh over pol > 0.001 on the FINAL U,
    n_iter CG iterations done (<= maxiter), n_field field evaluations spent."""
    det = lambda t: t.detach()
    positions, box, Q_local, pol, tholes = map(det, (positions, box, Q_local, pol, tholes))
    mScales, pScales, dScales = map(det, (mScales, pScales, dScales))
    n = positions.shape[0]
    U = torch.zeros(n, 3, dtype=torch.float64) if U_init is None else U_init.detach().clone()
    site = pol > 0.001
    Minv = (pol / DIELECTRIC)[:, None]

    def field(u):
        return force.grad_U_fn(positions, box, pairs, Q_local, u, pol, tholes, mScales, pScales, dScales)

    it, n_field = 0, 0
    while True:
        F = field(U)                                   # phase 0: true residual on the current U
        n_field += 1
        if torch.max(torch.abs(F[site])) < thresh:
            return U, True, it, n_field
        if it >= maxiter:
            return U, False, it, n_field
        r = -F
        z = Minv * r
        p = z.clone()
        rz = torch.sum(r * z)
        while True:                                    # phase 1: one field evaluation per iteration
            Ap = field(U + p) + r                      # field(U + p) - field(U), field(U) = -r
            n_field += 1
            pAp = torch.sum(p * Ap)
            if not (pAp > 0):                          # A not positive definite along p (polarization catastrophe)
                F = field(U)
                n_field += 1
                return U, False, it, n_field
            alpha = rz / pAp
            U = U + alpha * p
            r = r - alpha * Ap
            it += 1
            if torch.max(torch.abs(r[site])) < thresh or it >= maxiter:
                break                                  # back to phase 0: the true residual decides
            z = Minv * r
            rz_new = torch.sum(r * z)
            p = z + (rz_new / rz) * p
            rz = rz_new
