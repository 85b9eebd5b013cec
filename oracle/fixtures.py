"""Oracle (test infrastructure): harness inputs.

Loads the committed fixtures under tests/golden (made by tests/golden/make_fixtures.py
from the reference's shipped water1024.pdb / mpidwater.xml / dipole_1024 / ref_out)
and prepares the parameter set of examples/water_1024/run_admp.py:23-97 and
examples/water_pol_1024/run_admp.py:19-116: rc = 4 A, ethresh = 1e-4, forced
kappa = 0.657065221219616, lmax = 2, pmax = 10, m/p/dScales = [0,0,0,1,1].
Nothing here reads /root/reference.
"""
import os

import numpy as np
import torch

from .harmonics import cart2harm

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

KAPPA_EXAMPLE = 0.657065221219616     # run_admp.py:116
RC = 4.0
ETHRESH = 1e-4
SCALES = [0.0, 0.0, 0.0, 1.0, 1.0]

# hard-coded water parameters of examples/water_1024/run_admp.py:66-97 (O, H, H)
C6 = (37.19677405, 7.6111103, 7.6111103)
C8 = (85.26810658, 11.90220148, 11.90220148)
C10 = (134.44874488, 15.05074749, 15.05074749)
TT_Q = (-0.741706, 0.370853, 0.370853)
TT_B = (2.00095977, 1.999519942, 1.999519942)
TT_A = (458.3777, 0.0317, 0.0317)


class MoleculeCovalentMap:
    """Sparse stand-in for the reference's dense Na x Na ``covalent_map``
    (admp/parser.py:462-476): bonded partners per atom in CSR form.  ``lookup(i, j)``
    returns the bond count (0 when not bonded).  ``shape`` mimics the dense matrix."""

    def __init__(self, n_atoms, ci, cj, cn):
        order = np.lexsort((cj, ci))
        self.ci, self.cj, self.cn = ci[order].astype(np.int64), cj[order].astype(np.int64), cn[order].astype(np.int64)
        self.n_atoms = int(n_atoms)
        self.offsets = np.searchsorted(self.ci, np.arange(self.n_atoms + 1)).astype(np.int64)
        self._key = self.ci * self.n_atoms + self.cj

    @property
    def shape(self):
        return (self.n_atoms, self.n_atoms)

    def lookup(self, i, j):
        key = np.asarray(i, dtype=np.int64) * self.n_atoms + np.asarray(j, dtype=np.int64)
        pos = np.searchsorted(self._key, key)
        pos = np.minimum(pos, len(self._key) - 1)
        hit = self._key[pos] == key
        return np.where(hit, self.cn[pos], 0)

    def dense(self):
        m = np.zeros((self.n_atoms, self.n_atoms), dtype=np.int64)
        m[self.ci, self.cj] = self.cn
        return m

    def replicate(self, nrep):
        n = self.n_atoms
        ci = np.concatenate([self.ci + r * n for r in range(nrep)])
        cj = np.concatenate([self.cj + r * n for r in range(nrep)])
        cn = np.tile(self.cn, nrep)
        return MoleculeCovalentMap(n * nrep, ci, cj, cn)

    def subset(self, keep):
        """keep: sorted array of atom indices to retain (whole molecules)."""
        remap = -np.ones(self.n_atoms, dtype=np.int64)
        remap[keep] = np.arange(len(keep))
        sel = (remap[self.ci] >= 0) & (remap[self.cj] >= 0)
        return MoleculeCovalentMap(len(keep), remap[self.ci[sel]], remap[self.cj[sel]], self.cn[sel])


class WaterSystem:
    """A water box with the example scripts' parameters (float64 torch tensors)."""

    def __init__(self, positions, box_lengths, Q_cart, axis_type, axis_indices, pol, tholes, cov):
        self.positions = torch.as_tensor(np.ascontiguousarray(positions), dtype=torch.float64)
        self.box = torch.diag(torch.as_tensor(np.asarray(box_lengths), dtype=torch.float64))
        self.Q_local = cart2harm(torch.as_tensor(np.asarray(Q_cart), dtype=torch.float64), 2)
        self.Q_cart = np.asarray(Q_cart)
        self.axis_type = np.asarray(axis_type).astype(np.int64)
        self.axis_indices = np.asarray(axis_indices).astype(np.int64)
        self.pol = torch.as_tensor(np.asarray(pol), dtype=torch.float64)
        self.tholes = torch.as_tensor(np.asarray(tholes), dtype=torch.float64)
        self.covalent_map = cov
        self.n_atoms = self.positions.shape[0]
        self.mScales = torch.tensor(SCALES, dtype=torch.float64)
        self.pScales = torch.tensor(SCALES, dtype=torch.float64)
        self.dScales = torch.tensor(SCALES, dtype=torch.float64)
        nmol = self.n_atoms // 3
        self.c_list = torch.tensor(np.tile(np.array([C6, C8, C10]).T, (nmol, 1)), dtype=torch.float64)   # (Na,3)
        self.tt_a = torch.tensor(np.tile(TT_A, nmol), dtype=torch.float64)
        self.tt_b = torch.tensor(np.tile(TT_B, nmol), dtype=torch.float64)
        self.tt_q = torch.tensor(np.tile(TT_Q, nmol), dtype=torch.float64)

    def nonpol(self):
        """examples/water_1024/mpidwater.xml differs from the polarizable one only in
        the O polarizability (0.0)."""
        s = WaterSystem(self.positions.numpy(), torch.diagonal(self.box).numpy(), self.Q_cart, self.axis_type,
                        self.axis_indices, np.zeros(self.n_atoms), self.tholes.numpy(), self.covalent_map)
        return s

    def replicate(self, nx, ny, nz):
        """Replica-major replication."""
        L = torch.diagonal(self.box).numpy()
        pos, ai = [], []
        n = self.n_atoms
        r = 0
        for ix in range(nx):
            for iy in range(ny):
                for iz in range(nz):
                    pos.append(self.positions.numpy() + np.array([ix, iy, iz]) * L)
                    a = self.axis_indices.copy()
                    a[a >= 0] += r * n
                    ai.append(a)
                    r += 1
        nrep = nx * ny * nz
        return WaterSystem(np.concatenate(pos), L * np.array([nx, ny, nz]), np.tile(self.Q_cart, (nrep, 1)),
                           np.tile(self.axis_type, nrep), np.concatenate(ai), np.tile(self.pol.numpy(), nrep),
                           np.tile(self.tholes.numpy(), nrep), self.covalent_map.replicate(nrep))

    def carve(self, frac):
        """Sub-box [0, frac*L)^3 keeping whole molecules whose O lies inside; the box
        shrinks to frac*L (a small, still gas-like, periodic test system)."""
        L = torch.diagonal(self.box).numpy()
        pos = self.positions.numpy()
        o = pos[0::3]
        keep_mol = np.nonzero(np.all((o >= 0) & (o < frac * L), axis=1))[0]
        keep = (keep_mol[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        remap = -np.ones(self.n_atoms, dtype=np.int64)
        remap[keep] = np.arange(len(keep))
        ai = self.axis_indices[keep].copy()
        ai[ai >= 0] = remap[ai[ai >= 0]]
        return WaterSystem(pos[keep], L * frac, self.Q_cart[keep], self.axis_type[keep], ai,
                           self.pol.numpy()[keep], self.tholes.numpy()[keep], self.covalent_map.subset(keep))

    def jitter(self, seed, sigma=0.02):
        """Frame f of config C4: base + N(0, sigma) per coordinate, default_rng(seed)."""
        rng = np.random.default_rng(seed)
        return self.positions + torch.as_tensor(rng.normal(0.0, sigma, size=tuple(self.positions.shape)))


def _load(name):
    d = np.load(os.path.join(GOLDEN, name))
    n = d['positions'].shape[0]
    cov = MoleculeCovalentMap(n, d['cov_i'], d['cov_j'], d['cov_n'])
    s = WaterSystem(d['positions'], d['box'], d['Q_cart'], d['axis_type'], d['axis_indices'],
                    d['pol'], d['tholes'], cov)
    s.raw = d
    return s


def water1024():
    """examples/water_pol_1024 inputs (3072 atoms, 50 A cubic)."""
    return _load('water1024.npz')


def water2():
    """examples/water_pol_1024/water2.pdb (2 waters, 31.289 A cubic)."""
    return _load('water2.npz')


# rigid TIP3P-like geometry used for synthetic lattices (SURVEY 8(d) "dense" workload)
_R_OH = 0.9572
_ANGLE = 104.52 * np.pi / 180.0


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], 1),
                     np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], 1),
                     np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], 1)], 1)


def lattice_water(n_side, spacing=3.104, seed=7, jitter=0.1):
    """n_side^3 water molecules on a simple-cubic lattice (spacing in A), random orientation from
    default_rng(seed), centre jitter N(0, jitter); MPID water parameters of mpidwater.xml. A
    liquid-like system on which the reference's Jacobi SCF converges (the shipped 50 A box has
    O-O contacts of 1.14 A on which it diverges)."""
    base = water1024()
    rng = np.random.default_rng(seed)
    nmol = n_side ** 3
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing='ij'), -1).reshape(-1, 3).astype(np.float64)
    centres = (g + 0.5) * spacing + rng.normal(0.0, jitter, size=(nmol, 3))
    h = _ANGLE / 2
    local = np.array([[0.0, 0.0, 0.0],
                      [_R_OH * np.sin(h), 0.0, _R_OH * np.cos(h)],
                      [-_R_OH * np.sin(h), 0.0, _R_OH * np.cos(h)]])
    R = _random_rotations(rng, nmol)
    pos = (centres[:, None, :] + np.einsum('mab,kb->mka', R, local)).reshape(-1, 3)
    n = 3 * nmol
    mol = np.arange(nmol) * 3
    ai = np.empty((n, 3), dtype=np.int64)
    ai[0::3] = np.stack([mol + 1, mol + 2, -np.ones(nmol, dtype=np.int64)], 1)
    ai[1::3] = np.stack([mol, mol + 2, -np.ones(nmol, dtype=np.int64)], 1)
    ai[2::3] = np.stack([mol, mol + 1, -np.ones(nmol, dtype=np.int64)], 1)
    ci = np.concatenate([mol, mol, mol + 1, mol + 2, mol + 1, mol + 2])
    cj = np.concatenate([mol + 1, mol + 2, mol, mol, mol + 2, mol + 1])
    cn = np.concatenate([np.ones(4 * nmol), 2 * np.ones(2 * nmol)]).astype(np.int8)
    cov = MoleculeCovalentMap(n, ci, cj, cn)
    return WaterSystem(pos, np.full(3, n_side * spacing), np.tile(base.Q_cart[:3], (nmol, 1)),
                       np.tile(base.axis_type[:3], nmol), ai, np.tile(base.pol.numpy()[:3], nmol),
                       np.tile(base.tholes.numpy()[:3], nmol), cov)
