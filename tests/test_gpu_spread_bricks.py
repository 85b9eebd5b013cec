"""Brick-staged spread (admp_b200/csrc/spread_brick.cu; opt-in variant, see include/admp_b200.h) against the one-warp-per-atom scatter of the same library and
against a NumPy restatement of admp/recip.py:313-392 (charges), through the C ABI: partial edge bricks, stencils that wrap
around the periodic boundary in every dimension, atoms outside the cell, empty bricks, triclinic cells, both precisions,
charge-only (dispersion) and multipole + induced-dipole spreads, and bin reuse (admp_pme_spread_only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from admp_b200 import _lib                       # noqa: E402
from admp_b200._ctx import Context, to_dev       # noqa: E402

MESHES = [(48, 48, 48), (50, 67, 49), (154, 154, 154), (52, 48, 100), (64, 80, 97), (160, 48, 112)]


def _bspline6(u):
    """M6 on [0, 6) (admp/recip.py:80-101 restated as the Cox-de Boor recursion, float64)."""
    u = np.asarray(u, dtype=np.float64)

    def M(n, x):
        if n == 2:
            return np.where((x >= 0) & (x <= 2), 1.0 - np.abs(x - 1.0), 0.0)
        return x / (n - 1) * M(n - 1, x) + (n - x) / (n - 1) * M(n - 1, x - 1.0)
    return M(6, u)


def _numpy_charge_spread(pos, box, q, K):
    """Q(m) = sum_a q_a prod_d M6(u_ad - m_d) on the periodic mesh (admp/recip.py:313-392 for lmax = 0)."""
    K = np.asarray(K)
    inv = np.linalg.inv(box)
    nstar = (inv * K[None, :]).T                      # nstar[d] = K_d * inv[:, d]
    Q = np.zeros(tuple(K))
    for a in range(pos.shape[0]):
        x = nstar @ pos[a]
        m0 = np.ceil(x)
        f = m0 - x
        w = [_bspline6(f[d] + np.arange(6)) for d in range(3)]
        idx = [((m0[d] - 3 + np.arange(6)).astype(np.int64)) % K[d] for d in range(3)]
        # value at mesh point m0-3+k is M6(u - m) with u - m = (x + 3 - m0) ... = 3 - f - k + ... ; the kernel's convention:
        # weight k <-> M6(f + k) at mesh index m0 - 3 + k  (recip.cu mesh_anchor)
        Q[np.ix_(idx[0], idx[1], idx[2])] += q[a] * w[0][:, None, None] * w[1][None, :, None] * w[2][None, None, :]
    return Q


def _system(n, box, seed, spread_out=True):
    rng = np.random.default_rng(seed)
    L = np.abs(np.diag(box))
    pos = rng.uniform(-0.3, 1.3, size=(n, 3)) * L if spread_out else rng.uniform(0.0, 0.2, size=(n, 3)) * L
    # a few atoms pinned next to every face / corner so that their stencils wrap
    edge = np.array([[1e-3, 1e-3, 1e-3], [L[0] - 1e-3, L[1] - 1e-3, L[2] - 1e-3], [1e-3, L[1] - 1e-3, 0.5 * L[2]],
                     [0.5 * L[0], 1e-3, L[2] - 1e-3]])
    pos[:4] = edge
    M = rng.normal(size=(n, 10))
    U = 0.1 * rng.normal(size=(n, 3))
    return pos, M, U


def _spread(cx, K, pos, box, M, cols, stride, U, per_atom, only=False):
    sp, p = _lib.stream_ptr, _lib.ptr
    _lib.check(cx.lib.admp_ctx_set_spread(cx.handle, 0 if per_atom else 1))
    if only:
        _lib.check(cx.lib.admp_pme_spread_only(cx.handle, sp(), p(pos), p(M), cols, stride, p(U) if U is not None else None))
    else:
        _lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), cols, stride, p(U) if U is not None else None))
    torch.cuda.synchronize()
    return cx.mesh_view(K).clone()


@pytest.mark.parametrize('K', MESHES)
@pytest.mark.parametrize('precision', ['double', 'single'])
def test_brick_spread_equals_per_atom_spread(K, precision):
    cx = Context(precision)
    n = 700
    cx.set_pme(0.45, K[0], K[1], K[2], 2)
    cx.set_topology(n, None, None, None)
    assert cx.lib.admp_ctx_spread_bricks(cx.handle) == 0          # opt-in: the per-atom scatter is the default
    _lib.check(cx.lib.admp_ctx_set_spread(cx.handle, 1))
    assert cx.lib.admp_ctx_spread_bricks(cx.handle) in (16, 32)
    box_np = np.diag([23.0, 29.0, 31.0])
    pos_np, M_np, U_np = _system(n, box_np, seed=K[0] + K[2])
    dt = cx.dtype
    pos, box, M, U = (to_dev(torch.as_tensor(x), dt, cx.device) for x in (pos_np, box_np, M_np, U_np))
    tol = 1e-12 if dt == torch.float64 else 2e-5
    for cols, stride, Uarg in ((10, 10, U), (10, 10, None), (1, 10, None)):
        ref = _spread(cx, K, pos, box, M, cols, stride, Uarg, per_atom=True)
        got = _spread(cx, K, pos, box, M, cols, stride, Uarg, per_atom=False)
        scale = ref.abs().max().item()
        assert (got - ref).abs().max().item() <= tol * scale
        # untouched mesh points are exactly zero (the brick write-out is also the zero-fill)
        assert torch.equal(got == 0, ref == 0) or (got[ref == 0].abs().max().item() <= tol * scale)
    cx.close()


def test_brick_spread_charges_match_numpy_restatement():
    K = (50, 67, 49)
    cx = Context('double')
    n = 60
    cx.set_pme(0.45, K[0], K[1], K[2], 2)
    cx.set_topology(n, None, None, None)
    box_np = np.diag([13.0, 17.0, 11.0])
    pos_np, M_np, _ = _system(n, box_np, seed=5)
    pos, box, M = (to_dev(torch.as_tensor(x), cx.dtype, cx.device) for x in (pos_np, box_np, M_np))
    got = _spread(cx, K, pos, box, M, 1, 10, None, per_atom=False).cpu().numpy()
    ref = _numpy_charge_spread(pos_np, box_np, M_np[:, 0], K)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(got.sum() - M_np[:, 0].sum()) <= 1e-10 * np.abs(M_np[:, 0]).sum()      # partition of unity
    cx.close()


def test_brick_spread_triclinic_clustered_and_bin_reuse():
    K = (96, 64, 112)
    cx = Context('double')
    n = 1500
    cx.set_pme(0.45, K[0], K[1], K[2], 2)
    cx.set_topology(n, None, None, None)
    box_np = np.array([[20.0, 0.0, 0.0], [3.0, 18.0, 0.0], [-2.0, 4.0, 25.0]])
    # all atoms in one corner: most bricks are empty, a few hold hundreds of atoms (several staging batches per brick)
    pos_np, M_np, U_np = _system(n, np.diag([20.0, 18.0, 25.0]), seed=9, spread_out=False)
    pos, box, M, U = (to_dev(torch.as_tensor(x), cx.dtype, cx.device) for x in (pos_np, box_np, M_np, U_np))
    ref = _spread(cx, K, pos, box, M, 10, 10, U, per_atom=True)
    got = _spread(cx, K, pos, box, M, 10, 10, U, per_atom=False)
    assert (got - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
    # new dipoles on the same positions: the bins of the previous call are reused (the SCF cycles' case)
    U2 = to_dev(torch.as_tensor(-3.0 * U_np), cx.dtype, cx.device)
    ref2 = _spread(cx, K, pos, box, M, 10, 10, U2, per_atom=True)
    _spread(cx, K, pos, box, M, 10, 10, U, per_atom=False)
    got2 = _spread(cx, K, pos, box, M, 10, 10, U2, per_atom=False, only=True)
    assert (got2 - ref2).abs().max().item() <= 1e-12 * ref2.abs().max().item()
    assert (got2 - got).abs().max().item() > 1e-3 * ref.abs().max().item()
    cx.close()


def test_small_meshes_keep_the_per_atom_spread():
    cx = Context('double')
    cx.set_pme(0.45, 44, 42, 60, 2)
    cx.set_topology(8, None, None, None)
    _lib.check(cx.lib.admp_ctx_set_spread(cx.handle, 1))
    assert cx.lib.admp_ctx_spread_bricks(cx.handle) == 0
    cx.close()
