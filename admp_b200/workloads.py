"""Synthetic workloads of BASELINE.json (configs C1-C5 and the dense supplementary box), built
from the committed fixture of the reference's shipped water box (tests/golden/water1024.npz, made
by tests/golden/make_fixtures.py).  NumPy only; nothing here computes energies.

Parameters follow examples/water_1024/run_admp.py:23-97 and examples/water_pol_1024/run_admp.py:
19-116: rc = 4 A, ethresh = 1e-4, forced kappa = 0.657065221219616, K = 154 per 50 A, lmax = 2,
pmax = 10, m/p/dScales = [0,0,0,1,1].
"""
import os

import numpy as np

from .covalent import SparseCovalentMap
from .multipole import convert_cart2harm

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

KAPPA_EXAMPLE = 0.657065221219616
RC = 4.0
ETHRESH = 1e-4
K_BASE = 154
SCALES = np.array([0.0, 0.0, 0.0, 1.0, 1.0])
C6 = (37.19677405, 7.6111103, 7.6111103)
C8 = (85.26810658, 11.90220148, 11.90220148)
C10 = (134.44874488, 15.05074749, 15.05074749)


class Workload:
    """Plain container: positions (n,3), box (3,3), Q_local (n,9), axis tables, pol, tholes,
    covalent_map (SparseCovalentMap), c_list (n,3), scales, K (tuple) and kappa."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.n_atoms = self.positions.shape[0]


def _water_cov(nmol):
    mol = np.arange(nmol) * 3
    ci = np.concatenate([mol, mol, mol + 1, mol + 2, mol + 1, mol + 2])
    cj = np.concatenate([mol + 1, mol + 2, mol, mol, mol + 2, mol + 1])
    cn = np.concatenate([np.ones(4 * nmol), 2 * np.ones(2 * nmol)]).astype(np.int8)
    return SparseCovalentMap.from_pairs(3 * nmol, ci, cj, cn)


def _water_axes(nmol):
    mol = np.arange(nmol) * 3
    none = -np.ones(nmol, dtype=np.int64)
    ai = np.empty((3 * nmol, 3), dtype=np.int64)
    ai[0::3] = np.stack([mol + 1, mol + 2, none], 1)      # O : Bisector(z=H1, x=H2)
    ai[1::3] = np.stack([mol, mol + 2, none], 1)          # H1: ZThenX(z=O, x=H2)
    ai[2::3] = np.stack([mol, mol + 1, none], 1)
    return np.tile(np.array([1, 0, 0]), nmol), ai


def _assemble(pos, box_lengths, base, polarizable, reps=(1, 1, 1)):
    nmol = pos.shape[0] // 3
    at, ai = _water_axes(nmol)
    Qc = np.tile(base['Q_cart'][:3], (nmol, 1))
    pol = np.tile(base['pol'][:3], nmol) if polarizable else np.zeros(3 * nmol)
    th = np.tile(base['tholes'][:3], nmol)
    return Workload(positions=np.ascontiguousarray(pos), box=np.diag(np.asarray(box_lengths, dtype=np.float64)),
                    Q_local=convert_cart2harm(Qc, 2), axis_type=at, axis_indices=ai, pol=pol, tholes=th,
                    covalent_map=_water_cov(nmol), c_list=np.tile(np.array([C6, C8, C10]).T, (nmol, 1)),
                    mScales=SCALES.copy(), pScales=SCALES.copy(), dScales=SCALES.copy(),
                    K=tuple(K_BASE * r for r in reps), kappa=KAPPA_EXAMPLE, rc=RC, ethresh=ETHRESH, polarizable=polarizable)


def _base():
    return np.load(os.path.join(GOLDEN, 'water1024.npz'))


def water_box(reps=(1, 1, 1), polarizable=True):
    """The shipped 1024-water / 50 A box replicated reps = (nx, ny, nz) times, replica-major atom
    order, K = 154 * reps so the mesh spacing is unchanged: C1/C2 = (1,1,1),
    C3 = (2,4,4), C5 = (4,8,8)."""
    b = _base()
    L = b['box']
    pos = []
    for ix in range(reps[0]):
        for iy in range(reps[1]):
            for iz in range(reps[2]):
                pos.append(b['positions'] + np.array([ix, iy, iz]) * L)
    return _assemble(np.concatenate(pos), L * np.array(reps), b, polarizable, reps)


def jitter_frame(w, f, sigma=0.02):
    """Frame f of config C4: base positions + N(0, sigma) per coordinate, default_rng(1000 + f)."""
    rng = np.random.default_rng(1000 + f)
    return w.positions + rng.normal(0.0, sigma, size=w.positions.shape)


def dense_water(n_side=64, spacing=3.104, seed=7, jitter=0.1, polarizable=True):
    """Supplementary kernel-roofline input: n_side^3 rigid waters on a simple-cubic lattice at liquid
    density, random orientations from default_rng(seed) (SURVEY 8(d) "dense-256k" for n_side = 64)."""
    b = _base()
    rng = np.random.default_rng(seed)
    sides = (n_side,) * 3 if np.isscalar(n_side) else tuple(int(v) for v in n_side)      # cube, or (nx, ny, nz) molecules
    nmol = sides[0] * sides[1] * sides[2]
    g = np.stack(np.meshgrid(*[np.arange(v) for v in sides], indexing='ij'), -1).reshape(-1, 3).astype(np.float64)
    centres = (g + 0.5) * spacing + rng.normal(0.0, jitter, size=(nmol, 3))
    h = 104.52 * np.pi / 360.0
    local = np.array([[0.0, 0.0, 0.0], [0.9572 * np.sin(h), 0.0, 0.9572 * np.cos(h)], [-0.9572 * np.sin(h), 0.0, 0.9572 * np.cos(h)]])
    q = rng.normal(size=(nmol, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, bq, c, d = q.T
    R = np.stack([np.stack([a*a+bq*bq-c*c-d*d, 2*(bq*c-a*d), 2*(bq*d+a*c)], 1),
                  np.stack([2*(bq*c+a*d), a*a-bq*bq+c*c-d*d, 2*(c*d-a*bq)], 1),
                  np.stack([2*(bq*d-a*c), 2*(c*d+a*bq), a*a-bq*bq-c*c+d*d], 1)], 1)
    pos = (centres[:, None, :] + np.einsum('mab,kb->mka', R, local)).reshape(-1, 3)
    w = _assemble(pos, np.array(sides, dtype=np.float64) * spacing, b, polarizable)
    w.K = None
    return w
