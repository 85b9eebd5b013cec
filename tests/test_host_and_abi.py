"""CPU-side checks: the C-ABI library loads and exports every symbol include/admp_b200.h declares
(no compute calls without a GPU), the product path fails loudly without a device, and the host-side
helpers (covalent map, workloads, Ewald set-up) behave like the reference's."""
import os
import re

import numpy as np
import pytest
import torch

from admp_b200 import _lib, covalent, workloads
from admp_b200.multipole import convert_cart2harm
from admp_b200.pme import setup_ewald_parameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'admp_b200.h')).read()
    return sorted(set(re.findall(r'\b(admp_[a-z0-9_]+)\s*\(', hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), 'libadmp_b200.so does not export %s' % n
    assert set(_lib.EXPORTS) <= set(names)
    assert lib.admp_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_product_path_fails_loudly_without_gpu():
    from admp_b200.pme import ADMPPmeForce
    with pytest.raises(_lib.AdmpLibraryError):
        ADMPPmeForce(np.eye(3) * 50, np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3)), 4, 1e-4, 2)


def test_product_never_imports_the_oracle():
    """A product path that routes through the oracle voids every parity claim."""
    pkg = os.path.join(ROOT, 'admp_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn
            assert 'tools.analytic_proto' not in src


def test_sparse_covalent_map_matches_dense():
    dense = np.zeros((6, 6), dtype=int)
    for m in (0, 3):
        dense[m, m + 1] = dense[m + 1, m] = dense[m, m + 2] = dense[m + 2, m] = 1
        dense[m + 1, m + 2] = dense[m + 2, m + 1] = 2
    sp = covalent.as_sparse(dense)
    assert sp.shape == (6, 6)
    np.testing.assert_array_equal(sp.dense(), dense)
    np.testing.assert_array_equal(sp.offsets, [0, 2, 4, 6, 8, 10, 12])
    assert covalent.as_sparse(sp) is sp
    assert covalent.as_sparse(torch.tensor(dense)).index.tolist() == sp.index.tolist()


def test_setup_ewald_parameters_matches_reference_formula():
    kappa, K1, K2, K3 = setup_ewald_parameters(4, 1e-4, np.diag([50.0, 100.0, 25.0]))
    assert abs(kappa - np.sqrt(-np.log(2e-4)) / 4) < 1e-15
    assert (K1, K2, K3) == (154, 307, 77)


def test_workloads_follow_the_example_scripts():
    w = workloads.water_box((1, 1, 1))
    assert w.n_atoms == 3072 and w.K == (154, 154, 154) and w.kappa == 0.657065221219616
    np.testing.assert_allclose(w.Q_local[0], [-1.0614, -0.23671684, 0, 0, -0.0714102, 0, 0, 0.01106659, 0], rtol=1e-6, atol=1e-9)
    assert abs(w.pol[0] - 0.88) < 1e-6 and w.pol[1] == 0 and w.tholes[0] == 8.0
    np.testing.assert_array_equal(w.axis_type[:3], [1, 0, 0])
    np.testing.assert_array_equal(w.axis_indices[:3], [[1, 2, -1], [0, 2, -1], [0, 1, -1]])
    w2 = workloads.water_box((2, 1, 2))
    assert w2.n_atoms == 4 * 3072 and w2.K == (308, 154, 308)
    np.testing.assert_allclose(w2.positions[3072:6144], w.positions + np.array([0, 0, 50.0]))
    assert w2.axis_indices[3072, 0] == 3073
    f0, f1 = workloads.jitter_frame(w, 0), workloads.jitter_frame(w, 1)
    assert (f0 - w.positions).std() == pytest.approx(0.02, rel=0.05) and not np.allclose(f0, f1)
    np.testing.assert_array_equal(workloads.jitter_frame(w, 0), f0)


def test_convert_cart2harm_numpy_and_tensor():
    th = np.array([[-1.0614, 0, 0, -0.23671684, 0.0452889, 0.026121, -0.0714102, 0, 0, 0]])
    q = convert_cart2harm(th, 2)
    assert isinstance(q, np.ndarray) and q.shape == (1, 9)
    qt = convert_cart2harm(torch.tensor(th), 1)
    assert qt.shape == (1, 4) and float(qt[0, 1]) == pytest.approx(-0.23671684)
    with pytest.raises(NotImplementedError):
        convert_cart2harm(th, 3)


def test_neighbor_list_accepts_the_reference_scripts_call(monkeypatch):
    """examples/water_1024/run_admp.py:109-111: space.periodic_general(box) + partition.neighbor_list(displacement_fn,
    box, rc, 0, format=partition.OrderedSparse) keep working with `from admp_b200.neighbor import space, partition`."""
    import admp_b200.neighbor as nb
    calls = []
    monkeypatch.setattr(nb, 'NeighborListFn', lambda *a: calls.append(a))
    box = np.eye(3) * 50.0
    disp, shift = nb.space.periodic_general(box, fractional_coordinates=False)
    nb.partition.neighbor_list(disp, box, 4.0, 0, format=nb.partition.OrderedSparse)
    nb.neighbor_list(box, 4.0, 0.6)
    assert [(c[1], c[2], c[3]) for c in calls] == [(4.0, 0, 1.25), (4.0, 0.6, 1.25)] and all(c[0] is box for c in calls)
    d = disp(torch.tensor([[49.0, 0.0, 26.0]], dtype=torch.float64), torch.tensor([[1.0, 0.0, 1.0]], dtype=torch.float64))
    assert torch.allclose(d, torch.tensor([[-2.0, 0.0, -25.0]], dtype=torch.float64))      # +L/2 maps to -L/2 (A14)
