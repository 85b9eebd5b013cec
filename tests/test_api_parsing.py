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


def test_reference_own_forcefield_xml_parses_to_its_literals(tmp_path):
    """The reference's own input of the api.py front end (examples/openmm_api/forcefield.xml, committed as a data fixture
    by tests/golden/make_fixtures.py) on the topology of its water1024.pdb: the generators hold the file's literals, the
    residue template is matched by atom NAME (the template lists H1, H2, O; the PDB O, H1, H2), anchors follow
    admp/api.py:44-116."""
    import os
    ff = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'openmm_api_forcefield.xml')
    s = fixtures.water1024()
    _, pdbfile = _write_inputs(tmp_path, s)
    H = api.Hamiltonian(ff)
    pdb = api.PDBFile(pdbfile)
    disp_gen, pme_gen = H.getGenerators()
    assert pme_gen.lmax == 2 and pme_gen.pmax == 10 and pme_gen.lpol and disp_gen.pmax == 10
    assert pme_gen.types.tolist() == ['380', '381'] and disp_gen.types.tolist() == ['380', '381']
    ip = pme_gen._input_params
    assert ip['c0'].tolist() == [-1.0614, 0.5307] and ip['dZ'].tolist() == [-0.023671684, 0.0]
    assert ip['qXX'].tolist() == [0.000150963, 0.0] and ip['qYY'].tolist() == [0.00008707, 0.0] and ip['qZZ'].tolist() == [-0.000238034, 0.0]
    assert ip['polarizabilityXX'].tolist() == [0.00088, 0.0] and ip['thole'].tolist() == [8.0, 0.0]
    assert pme_gen.kStrings == {'kz': ['-381', '380'], 'kx': ['-381', '381'], 'ky': ['', '']}
    for k in ('mScales', 'pScales', 'dScales'):
        assert pme_gen._scales[k] == [0.0, 0.0, 0.0, 1.0, 1.0]
    raw = disp_gen._raw
    assert raw['A'].tolist() == [1203470.743, 83.2283563] and raw['B'].tolist() == [37.81265679, 37.78544799]
    assert raw['Q'].tolist() == [-0.741706, 0.370853] and raw['C6'].tolist() == [0.001383816, 5.7929e-05]
    assert raw['C8'].tolist() == [7.27065e-05, 1.416624e-06] and raw['C10'].tolist() == [1.8076465e-6, 2.26525e-08]
    assert raw['mScales'].tolist() == [0.0, 0.0, 0.0, 1.0, 1.0]
    types, bonds = H._match(pdb.topology)
    assert [types[i] for i in range(6)] == ['380', '381', '381'] * 2 and len(bonds) == 2 * 1024
    data = api._Data(pdb.topology, types, bonds)
    m = np.array([int(np.where(pme_gen.types == data.atomType[a])[0][0]) for a in data.atoms])
    at, names = api.set_axis_type(m, pme_gen.types, pme_gen.kStrings)
    ai = api._map_axis_indices(data, names)
    assert at.tolist() == s.axis_type.tolist() and np.array_equal(ai[:, :2], s.axis_indices[:, :2])
    cov = api.build_covalent_map(data, 6).dense()
    assert np.array_equal(cov[:3, :3], [[0, 1, 1], [1, 0, 2], [1, 2, 0]]) and cov[0, 3:].sum() == 0


def test_quasi_internal_frame_and_induced_dipole_rotation_helpers_reference_literals():
    """The reference's own known answers for build_quasi_internal (tests/test_sptial.py:11-41); rot_ind_global2local
    (admp/multipole.py:80-89) against the definition R[zxy][:, zxy] . U on the same frames. CPU-only helpers."""
    import torch
    from admp_b200.spatial import build_quasi_internal
    r1, r2 = np.array([[0.0, 0, 0], [0.0, 0, 0]]), np.array([[1.0, 0, 0], [1.0, 1, 0]])
    dr, nrm = r2.copy(), np.array([1.0, 1.414213])
    fr = build_quasi_internal(r1, r2, dr, nrm).numpy()
    np.testing.assert_allclose(fr[0], [[0.0, 1.0, 0.0], [0, 0, 1], [1, 0, 0]], atol=1e-12)
    np.testing.assert_allclose(fr[1], [[0.70710534, -0.70710814, 0.0], [0.0, 0.0, -1.0000004], [0.70710707, 0.70710707, 0.0]],
                               rtol=1e-5, atol=3e-6)      # float32-era literals, norm given to 7 digits
    # rot_ind_global2local needs the CUDA library (admp_rotate); its definition is checked here on the oracle twin and on
    # the GPU in tests/test_gpu_parity.py::test_frames_and_rotation_match_reference_literals
    from oracle.harmonics import rot_ind_global2local as o_rot
    U = torch.tensor([[0.3, -0.2, 0.5], [0.1, 0.4, -0.7]], dtype=torch.float64)
    R = torch.tensor(fr)
    got = o_rot(U, R)
    zxy = [2, 0, 1]
    want = torch.stack([R[k][zxy][:, zxy] @ U[k] for k in range(2)])
    assert torch.allclose(got, want, atol=1e-14)
