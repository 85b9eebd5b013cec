"""The two real-space pair traversals (flat rows / j-cluster x i-lane tiles, admp_b200/csrc/pair_cluster.cu) must give the
same energies and gradients: both against the reference-source goldens (1e-6) and against each other (summation order
only: 1e-10), for every output the pair pass produces; lists the cluster path cannot take (unsorted rows) must fall back."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import refcases, fixtures, pairlist                 # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def dev(x, grad=True, dtype=torch.float64):
    return torch.tensor(np.asarray(x), device='cuda', dtype=dtype, requires_grad=grad)


def force_path(calc, force):
    from admp_b200 import _lib
    _lib.check(calc._ctx.lib.admp_ctx_set_pair_cluster(calc._ctx.handle, force, 0))


def active(calc):
    return calc._ctx.lib.admp_ctx_pair_cluster_active(calc._ctx.handle)


@pytest.mark.parametrize('name', ['lattice4', 'carved'])
def test_cluster_and_flat_traversals_agree_with_the_reference_and_each_other(name):
    from admp_b200.pme import ADMPPmeForce
    c = refcases.get(name)
    g = np.load(os.path.join(GOLDEN, 'ref_%s.npz' % name))
    s = c.s
    out = {}
    for force in (-1, 1):
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, c.rc, c.ethresh, 2, lpol=True)
        force_path(calc, force)
        p, b, q, u = dev(s.positions), dev(s.box), dev(c.Q_pert), dev(c.U_pert)
        pol, th, m, ps = dev(c.pol_pert), dev(c.tholes_pert), dev(c.mScales_pert), dev([0.0, 0.4, 0.0, 1.0, 1.0])
        E = calc.energy_fn(p, b, c.pairs, q, u, pol, th, m, ps, s.dScales)
        grads = torch.autograd.grad(E, [p, b, q, u, pol, th, m, ps])
        assert active(calc) == (1 if force > 0 else 0)
        # SCF (field-only kernel) + wrapped energy with the water parameters
        E2, F2, dbox2 = calc.get_forces_and_virial(s.positions, s.box, c.pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        assert active(calc) == (1 if force > 0 else 0)
        out[force] = (E, grads, E2, F2, dbox2, calc.U_ind.clone(), calc.n_cycle)
        # non-polarizable kernels
        calc0 = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, c.rc, c.ethresh, 2)
        force_path(calc0, force)
        p0, q0, m0 = dev(s.positions), dev(c.Q_pert), dev(c.mScales_pert)
        E0 = calc0.get_energy(p0, s.box, c.pairs, q0, m0)
        g0 = torch.autograd.grad(E0, [p0, q0, m0])
        assert rel(E0, g['nonpol_E']) < 1e-6 and rel(g0[0], g['nonpol_dpos']) < 1e-6 and rel(g0[1], g['nonpol_dQ']) < 1e-6
        assert rel(g0[2], g['nonpol_dmScales']) < 1e-6
    for force in (-1, 1):
        E, grads, E2, F2, dbox2, U, ncyc = out[force]
        # energy_fn at pScales = [0, .4, 0, 1, 1] is not in the goldens (they use the water pScales): compare the SCF run
        assert ncyc == int(g['scf_n_cycle'])
        assert rel(E2, g['scf_E']) < 1e-6 and rel(F2, g['scf_dpos']) < 1e-6 and rel(U, g['scf_U']) < 1e-6
    a, b = out[-1], out[1]
    assert rel(a[0], b[0]) < 1e-10
    for x, y in zip(a[1], b[1]):
        assert rel(x, y) < 1e-9
    assert rel(a[2], b[2]) < 1e-10 and rel(a[3], b[3]) < 1e-9 and rel(a[4], b[4]) < 1e-9 and rel(a[5], b[5]) < 1e-9


def test_unsorted_rows_fall_back_to_the_flat_kernel_and_dense_lists_select_the_cluster_kernel():
    from admp_b200 import _lib   # noqa: F401
    from admp_b200.pme import ADMPPmeForce
    s = fixtures.lattice_water(6, 3.1, seed=3)                    # 216 waters, 18.6 A box
    rc = 7.0
    pairs, n = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), rc)
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, rc, 1e-4, 2, lpol=True)
    args = (s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    # oracle / admp_nblist_build order: rows sorted by (i, j) -> clusters on column 0; dense list -> the device picks the cluster kernel
    E_a, F_a = calc.get_forces(s.positions, s.box, pairs, *args)
    assert active(calc) == 1, 'dense sorted list should select the cluster kernel (%d rows, %d clusters)' % (n, s.n_atoms // 3)
    # grouped by j, ascending i (jax_md OrderedSparse-like order) with padding rows -> clusters on column 1
    order = np.lexsort((pairs[:n, 0], pairs[:n, 1]))
    sorted_pairs = np.concatenate([pairs[:n][order], np.full((37, 2), s.n_atoms, dtype=pairs.dtype)])
    E_b, F_b = calc.get_forces(s.positions, s.box, sorted_pairs, *args)
    assert active(calc) == 1
    assert rel(E_a, E_b) < 1e-10 and rel(F_a, F_b) < 1e-9
    # shuffled rows: no usable order -> flat kernel, same result
    rng = np.random.default_rng(0)
    shuffled = pairs[:n][rng.permutation(n)]
    E_s, F_s = calc.get_forces(s.positions, s.box, shuffled, *args)
    assert active(calc) == 0
    assert rel(E_s, E_b) < 1e-10 and rel(F_s, F_b) < 1e-9
    force_path(calc, -1)
    E_f, F_f = calc.get_forces(s.positions, s.box, pairs, *args)
    assert active(calc) == 0 and rel(E_f, E_a) < 1e-10 and rel(F_f, F_a) < 1e-9
    force_path(calc, 0)
    # the same list from the library's own neighbour list
    from admp_b200.neighbor import neighbor_list
    nbr = neighbor_list(s.box, rc).allocate(s.positions)
    E_c, F_c = calc.get_forces(s.positions, s.box, nbr.pairs, *args)
    assert active(calc) == 1 and rel(E_c, E_b) < 1e-12 and rel(F_c, F_b) < 1e-10
    # a reversed row (i > j) in the middle is "not evaluated" (pme.py:671) and breaks the prefix rule -> flat
    broken = sorted_pairs.copy()
    broken[5] = broken[5][::-1]
    E_d, F_d = calc.get_forces(s.positions, s.box, broken, *args)
    assert active(calc) == 0
    ref = np.delete(sorted_pairs, 5, axis=0)
    E_e, F_e = calc.get_forces(s.positions, s.box, ref, *args)
    assert rel(E_d, E_e) < 1e-10 and rel(F_d, F_e) < 1e-9


def test_cluster_kernel_single_precision():
    from admp_b200 import settings
    from admp_b200.pme import ADMPPmeForce
    c = refcases.get('lattice4')
    g = np.load(os.path.join(GOLDEN, 'ref_lattice4.npz'))
    s = c.s
    old = settings.PRECISION
    settings.PRECISION = 'single'
    try:
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, c.rc, c.ethresh, 2, lpol=True)
        force_path(calc, 1)
        E, F = calc.get_forces(s.positions, s.box, c.pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        assert active(calc) == 1
    finally:
        settings.PRECISION = old
    assert calc.n_cycle == int(g['scf_n_cycle'])
    assert rel(E, g['scf_E']) < 1e-4 and rel(F, g['scf_dpos']) < 1e-4


# ------------------------------------------------------------------------------------------ generic pair framework
def test_generic_pair_kernel_any_callable_and_fused_variants():
    """generate_pairwise_interaction (admp/pairwise.py:45-91) accepts ANY per-pair kernel: a user callable runs between
    the CUDA geometry kernel and its adjoint; the fused TT kernels agree with the same formula evaluated that way, with
    the oracle's autograd, and (c6 variant) with the reference-source goldens."""
    from admp_b200.pairwise import (generate_pairwise_interaction, TT_damping_qq_c6_kernel, TT_damping_qq_c6_c8_c10_kernel,
                                    PAIR_KERNELS)
    from oracle.shortrange import generate_pairwise_interaction as o_gen
    assert set(PAIR_KERNELS) == {'TT_damping_qq_c6_kernel', 'TT_damping_qq_c6_c8_c10_kernel'}
    c = refcases.get('carved')
    g = np.load(os.path.join(GOLDEN, 'ref_carved.npz'))
    s = c.s
    vals6 = [s.positions, s.box, c.mScales_pert, s.tt_a, s.tt_b, s.tt_q, s.c_list[:, 0]]

    def run(kernel, vals):
        fn = generate_pairwise_interaction(kernel, s.covalent_map, static_args={})
        t = [dev(v) for v in vals]
        E = fn(t[0], t[1], c.pairs, t[2], *t[3:])
        return E, torch.autograd.grad(E, t)

    # 1. the reference's kernel: fused CUDA body == the same formula through the generic path == reference source
    E_f, g_f = run(TT_damping_qq_c6_kernel, vals6)
    E_g, g_g = run(lambda dr, m, *pp: TT_damping_qq_c6_kernel(dr, m, *pp), vals6)
    assert rel(E_f, g['tt_E']) < 1e-6 and rel(E_g, g['tt_E']) < 1e-6
    for a, b, key in zip(g_g, g_f, ('pos', 'box', 'mS', 'a', 'b', 'q', 'c')):
        assert rel(a, b) < 1e-8, key
    assert rel(g_g[0], g['tt_dpos']) < 1e-6 and rel(g_g[2], g['tt_dmScales']) < 1e-6 and rel(g_g[3], g['tt_da']) < 1e-6

    # 2. a kernel the library has never seen (Buckingham + screened charge): generic path vs the oracle driver + autograd
    def my_kernel(dr, m, Ai, Aj, Bi, Bj, qi, qj):
        return m * (torch.sqrt(Ai * Aj) * torch.exp(-0.5 * (Bi + Bj) * dr) + 138.935 * qi * qj * torch.erfc(0.3 * dr) / dr)
    vals = [s.positions, s.box, c.mScales_pert, s.tt_a, s.tt_b, s.tt_q]
    E_u, g_u = run(my_kernel, vals)
    to = [torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for v in vals]
    E_o = o_gen(my_kernel, s.covalent_map, {})(to[0], to[1], c.pairs, to[2], *to[3:])
    g_o = torch.autograd.grad(E_o, to)
    assert rel(E_u, E_o) < 1e-10
    for a, b, key in zip(g_u, g_o, ('pos', 'box', 'mS', 'A', 'B', 'q')):
        if key == 'box':
            assert rel(torch.diagonal(a), torch.diagonal(b)) < 1e-9
        else:
            assert rel(a, b) < 1e-9, key

    # 3. the C8 / C10 extension: fused CUDA body vs its formula through the oracle driver
    vals10 = vals6[:6] + [s.c_list[:, 0], s.c_list[:, 1], s.c_list[:, 2]]
    E_10, g_10 = run(TT_damping_qq_c6_c8_c10_kernel, vals10)
    to = [torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for v in vals10]
    E_o = o_gen(lambda dr, m, *pp: TT_damping_qq_c6_c8_c10_kernel(dr, m, *pp), s.covalent_map, {})(to[0], to[1], c.pairs, to[2], *to[3:])
    g_o = torch.autograd.grad(E_o, to)
    assert rel(E_10, E_o) < 1e-9
    for k, (a, b) in enumerate(zip(g_10, g_o)):
        if k == 1:
            assert rel(torch.diagonal(a), torch.diagonal(b)) < 1e-8
        else:
            assert rel(a, b) < 1e-8, k
    # padding / reversed rows contribute nothing in the generic path either
    fn = generate_pairwise_interaction(my_kernel, s.covalent_map, {})
    padded = np.concatenate([c.pairs[:c.n_pairs], np.full((11, 2), s.n_atoms, dtype=c.pairs.dtype), c.pairs[:3, ::-1]])
    assert rel(fn(s.positions, s.box, padded, c.mScales_pert, s.tt_a, s.tt_b, s.tt_q), E_u) < 1e-12
    assert fn(s.positions, s.box, np.zeros((0, 2), dtype=np.int32), c.mScales_pert, s.tt_a, s.tt_b, s.tt_q).item() == 0.0


def test_two_mesh_scf_body_matches_single_mesh(monkeypatch):
    """Meshes beyond L2 run the SCF cycles with a second real buffer (zero-fill of the spread mesh overlapped with the
    spectrum passes, api.cu scf_body); forced on here for a small mesh: same cycles, same result as the one-mesh body."""
    from admp_b200.pme import ADMPPmeForce
    c = refcases.get('lattice4')
    g = np.load(os.path.join(GOLDEN, 'ref_lattice4.npz'))
    s = c.s
    args = (c.pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    out = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('ADMP_TWO_MESH', mode)
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, c.rc, c.ethresh, 2, lpol=True)
        E, F, V = calc.get_forces_and_virial(s.positions, s.box, *args)
        E2, F2 = calc.get_forces(s.positions, s.box, *args)                 # no virial pass: the final gather reads the second buffer
        assert calc.n_cycle == int(g['scf_n_cycle'])
        assert rel(E, g['scf_E']) < 1e-6 and rel(F, g['scf_dpos']) < 1e-6 and rel(E2, g['scf_E']) < 1e-6 and rel(F2, g['scf_dpos']) < 1e-6
        out[mode] = (E, F, V, calc.U_ind.clone())
    for a, b in zip(out['0'], out['1']):
        assert rel(a, b) < 1e-10
