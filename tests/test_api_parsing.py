"""CPU part of the force-field front end (admp_b200/api.py): XML / PDB parsing, residue-template matching, the
AMOEBA-style anchor convention (admp/api.py:44-116) and the sparse covalent map (admp/api.py:24-42). No GPU calls."""
import numpy as np

from admp_b200 import api
from oracle import fixtures
from test_gpu_api import _write_inputs


def test_parsing_matching_axes_and_covalent_map(tmp_path):
    s = fixtures.lattice_water(3, 3.3, seed=2)
    ff, pdbfile = _write_inputs(tmp_path, s)
    H = api.Hamiltonian(ff)
    pdb = api.PDBFile(pdbfile)
    assert pdb.topology.getNumAtoms() == s.n_atoms
    assert np.allclose(pdb.topology.box, s.box.numpy()) and np.allclose(pdb.positions, s.positions.numpy(), atol=5.1e-4)
    disp_gen, pme_gen = H.getGenerators()
    assert isinstance(disp_gen, api.ADMPDispGenerator) and isinstance(pme_gen, api.ADMPPmeGenerator)
    assert pme_gen.lmax == 2 and pme_gen.lpol and disp_gen.pmax == 10
    types, bonds = H._match(pdb.topology)
    assert [types[i] for i in range(3)] == ['380', '381', '381'] and len(bonds) == 2 * (s.n_atoms // 3)
    data = api._Data(pdb.topology, types, bonds)
    cov = api.build_covalent_map(data, 6)
    dense = cov.dense()
    assert np.array_equal(dense[:3, :3], [[0, 1, 1], [1, 0, 2], [1, 2, 0]]) and dense[0, 3:].sum() == 0
    m = np.array([int(np.where(pme_gen.types == data.atomType[a])[0][0]) for a in data.atoms])
    at, names = api.set_axis_type(m, pme_gen.types, pme_gen.kStrings)
    ai = api._map_axis_indices(data, names)
    assert at.tolist() == s.axis_type.tolist()                 # O: Bisector, H: ZThenX
    assert np.array_equal(ai[:, :2], s.axis_indices[:, :2]) and (ai[:, 2] == -1).all()


def test_axis_type_table():
    """every branch of the anchor convention (admp/api.py:98-112)"""
    types = np.array(['a', 'b', 'c', 'd', 'e', 'f'])
    k = {'kz': ['', 'b', '-b', 'b', '-b', 'b'], 'kx': ['', '', '-c', 'c', '-c', '-c'], 'ky': ['', '', '', '', '-d', '-d']}
    at, names = api.set_axis_type(range(6), types, k)
    assert at.tolist() == [api.NoAxisType, api.Zonly, api.Bisector, api.ZThenX, api.ThreeFold, api.ZBisect]
    assert names[4] == ['e', 'b', 'c', 'd']
