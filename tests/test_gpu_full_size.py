"""Parity at BASELINE.json's full sizes through a size-independent property: the replicated boxes C3
(32k waters, mesh 308x616x616) and C5 (256k waters, mesh 616x1232x1232) are exact periodic replicas of the base cell with
the mesh spacing kept, so E = n_rep * E(C2), and forces and induced dipoles of every replica equal the base cell's - at sizes
the oracle (and the reference) cannot run. The base cell itself is checked against the oracle elsewhere (test_gpu_parity)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def _rel(a, b):
    return (a - b).abs().max().item() / b.abs().max().item()


@pytest.fixture(scope='module')
def base():
    import run_config
    return run_config.run('C2', 1, disp=True)


def test_c3_replica_invariance_polarizable_and_dispersion(base):
    import run_config
    res = run_config.run('C3', 1, disp=True)
    n = res['nrep']
    assert n == 32
    assert abs(res['E'] / (n * base['E']) - 1) < 1e-10
    assert _rel(res['F'], base['F']) < 1e-9 and _rel(res['U'], base['U']) < 1e-9
    assert abs(res['Ed'] / (n * base['Ed']) - 1) < 1e-10


def test_c5_replica_invariance_polarizable(base):
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip('needs ~20 GB of free device memory')
    import run_config
    res = run_config.run('C5', 1)
    n = res['nrep']
    assert n == 256
    assert abs(res['E'] / (n * base['E']) - 1) < 1e-10
    assert _rel(res['F'], base['F']) < 1e-8 and _rel(res['U'], base['U']) < 1e-9
