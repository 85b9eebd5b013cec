"""Oracle (test infrastructure): the fixed inputs of the reference-source golden vectors.

One definition shared by tests/golden/make_reference_goldens.py (which feeds them to the UNMODIFIED
reference sources under oracle/jaxshim and writes tests/golden/ref_*.npz) and by the tests that
compare the oracle (CPU) and the CUDA path (GPU) with those files.  Cubic boxes only: for K1=K2=K3
and a cubic cell the reference's k-vector ordering and transposed spline Jacobian
coincide with the chain-rule-correct forms.
"""
import numpy as np
import torch

from . import fixtures, pairlist

SMALL = ('lattice3', 'lattice4', 'carved')
FULL = ('c1', 'c2')


class Case:
    pass


def _perturbed(s, seed=3):
    """Generic (non-water-like) parameters so that every term of the kernels is exercised."""
    rng = np.random.default_rng(seed)
    Ql = s.Q_local.numpy().copy()
    Ql[:, 1:] += rng.normal(0, 0.05, (s.n_atoms, 8))
    pol = np.abs(rng.normal(0.8, 0.2, s.n_atoms))
    pol[1::7] = 0.0
    th = np.abs(rng.normal(3.0, 1.0, s.n_atoms))
    U = rng.normal(0, 0.05, (s.n_atoms, 3))
    f64 = lambda a: torch.tensor(a, dtype=torch.float64)
    return f64(Ql), f64(U), f64(pol), f64(th)


def get(name):
    c = Case()
    c.name = name
    c.ethresh = 1e-4
    c.kappa = None            # None: keep setup_ewald_parameters' value
    if name == 'lattice3':    # 27 waters, 9.6 A, liquid-like: Jacobi SCF converges in 2 cycles
        s = fixtures.lattice_water(3, 3.2, seed=11)
        c.rc = 4.0
    elif name == 'lattice4':  # 64 waters, 12.6 A, liquid-like
        s = fixtures.lattice_water(4, 3.15, seed=5)
        c.rc = 5.0
    elif name == 'carved':    # 99 waters carved from the shipped box, 25 A, gas-like (SCF diverges)
        s = fixtures.water1024().carve(0.5)
        c.rc = 6.0
    elif name == 'c1':        # examples/water_1024 (BASELINE config 0)
        s = fixtures.water1024().nonpol()
        c.rc, c.kappa = fixtures.RC, fixtures.KAPPA_EXAMPLE
    elif name == 'c2':        # examples/water_pol_1024 (BASELINE config 1, the headline)
        s = fixtures.water1024()
        c.rc, c.kappa = fixtures.RC, fixtures.KAPPA_EXAMPLE
    else:
        raise KeyError(name)
    c.s = s
    c.pairs, c.n_pairs = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), c.rc)
    if name in SMALL:
        c.Q_pert, c.U_pert, c.pol_pert, c.tholes_pert = _perturbed(s)
        rng = np.random.default_rng(17)
        c.mScales_pert = torch.tensor([0.0, 0.3, 0.6, 0.9, 1.0], dtype=torch.float64)
        c.c_list_pert = s.c_list * torch.tensor(rng.uniform(0.8, 1.2, tuple(s.c_list.shape)))
    return c
