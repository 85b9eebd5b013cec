"""Hand-written mixed-radix FFT (admp_b200/csrc/fft.cu) against torch.fft and against the cuFFT
backend of the same library; fused convolution round trip against the unfused path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from admp_b200 import _lib                       # noqa: E402
from admp_b200._ctx import Context               # noqa: E402

SIZES = [(154, 154, 154), (44, 42, 60), (22, 26, 30), (6, 10, 14), (308, 154, 22), (96, 100, 98),
         (616, 22, 308), (22, 1232, 616), (26, 308, 1232)]


def _ctx(K, precision='double'):
    cx = Context(precision)
    cx.set_pme(0.45, K[0], K[1], K[2], 2)
    return cx


@pytest.mark.parametrize('K', SIZES)
@pytest.mark.parametrize('precision', ['double', 'single'])
def test_forward_and_inverse_match_torch_fft(K, precision):
    cx = _ctx(K, precision)
    assert cx.lib.admp_ctx_fft_backend(cx.handle) == 1
    dt = cx.dtype
    cdt = torch.complex128 if dt == torch.float64 else torch.complex64
    g = torch.Generator(device='cuda').manual_seed(K[0] * 7 + K[2])
    mesh = torch.randn(K, dtype=dt, device='cuda', generator=g)
    sp, p = _lib.stream_ptr, _lib.ptr
    _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(mesh), mesh.numel() * mesh.element_size(), 1))
    _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 0))
    spec = torch.empty((K[0], K[1], K[2] // 2 + 1), dtype=cdt, device='cuda')
    _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 1, p(spec), spec.numel() * spec.element_size(), 0))
    ref = torch.fft.rfftn(mesh.double())
    tol = 1e-12 if dt == torch.float64 else 2e-5
    assert (spec.to(torch.complex128) - ref).abs().max().item() < tol * ref.abs().max().item()
    _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 1))
    back = torch.empty_like(mesh)
    _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(back), back.numel() * back.element_size(), 0))
    n = K[0] * K[1] * K[2]
    assert (back.double() / n - mesh.double()).abs().max().item() < tol * 10


@pytest.mark.parametrize('K', [(154, 154, 154), (44, 42, 60)])
@pytest.mark.parametrize('kind', [_lib.CK_COULOMB, _lib.CK_DISP6, _lib.CK_DISP10])
def test_fused_roundtrip_equals_cufft_plus_convolve(K, kind):
    cx = _ctx(K)
    cx.set_topology(4, None, None, None)
    sp, p = _lib.stream_ptr, _lib.ptr
    box = torch.diag(torch.tensor([31.0, 29.0, 37.0], dtype=torch.float64, device='cuda'))
    pos = torch.rand((4, 3), dtype=torch.float64, device='cuda') * 20
    Q = torch.rand((4, 1), dtype=torch.float64, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(3)
    mesh = torch.randn(K, dtype=torch.float64, device='cuda', generator=g)
    out = {}
    for backend in (1, 0):
        _lib.check(cx.lib.admp_ctx_set_fft_backend(cx.handle, backend))
        # admp_pme_spread sets up the box on the context; its mesh is then overwritten
        _lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(Q), 1, 1, None))
        _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(mesh), mesh.numel() * 8, 1))
        scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device='cuda')
        if backend == 1:
            _lib.check(cx.lib.admp_pme_fft_convolve(cx.handle, sp(), kind, _lib.WANT_VIRIAL, p(scal)))
        else:
            _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 0))
            _lib.check(cx.lib.admp_pme_convolve(cx.handle, sp(), kind, _lib.WANT_VIRIAL, p(scal)))
            _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 1))
        phi = torch.empty_like(mesh)
        _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(phi), phi.numel() * 8, 0))
        out[backend] = (phi, scal.clone())
    a, b = out[1], out[0]
    assert (a[0] - b[0]).abs().max().item() < 1e-11 * b[0].abs().max().item()
    assert abs(a[1][_lib.S_E_RECIP].item() - b[1][_lib.S_E_RECIP].item()) < 1e-11 * abs(b[1][_lib.S_E_RECIP].item())
    tk_a, tk_b = a[1][_lib.S_TK:_lib.S_TK + 6], b[1][_lib.S_TK:_lib.S_TK + 6]
    assert (tk_a - tk_b).abs().max().item() < 1e-10 * tk_b.abs().max().item()


def test_unsupported_sizes_fall_back_to_cufft():
    cx = _ctx((31, 62, 34))          # 31 and 17 are outside the radix set
    assert cx.lib.admp_ctx_fft_backend(cx.handle) == 0
    assert cx.lib.admp_ctx_set_fft_backend(cx.handle, 1) != 0
    cx2 = _ctx((22, 22, 21))         # odd K3
    assert cx2.lib.admp_ctx_fft_backend(cx2.handle) == 0


def test_full_evaluation_identical_on_both_backends():
    """Config C1 (154^3): energy, forces and virial with the fused hand-written FFT vs cuFFT."""
    from admp_b200 import workloads
    from admp_b200.pme import ADMPPmeForce
    from admp_b200.neighbor import neighbor_list
    w = workloads.water_box((1, 1, 1), polarizable=True)
    pairs = neighbor_list(w.box, w.rc).allocate(w.positions).pairs
    res = {}
    for backend in (1, 0):
        calc = ADMPPmeForce(w.box, w.axis_type, w.axis_indices, w.covalent_map, w.rc, w.ethresh, 2, lpol=True)
        calc.update_env('kappa', w.kappa)
        _lib.check(calc._ctx.lib.admp_ctx_set_fft_backend(calc._ctx.handle, backend))
        E, F, V = calc.get_forces_and_virial(w.positions, w.box, pairs, w.Q_local, w.pol, w.tholes, w.mScales, w.pScales)
        res[backend] = (E.item(), F.clone(), V.clone(), calc.n_cycle)
    assert res[0][3] == res[1][3]
    assert abs(res[0][0] - res[1][0]) < 1e-9 * abs(res[0][0])
    assert (res[0][1] - res[1][1]).abs().max().item() < 1e-9 * res[0][1].abs().max().item()
    assert (res[0][2] - res[1][2]).abs().max().item() < 1e-8 * res[0][2].abs().max().item()


@pytest.mark.parametrize('env', [{'ADMP_FFT_WIDE': '0'}, {'ADMP_FFT_WIDE': '1'}, {'ADMP_FFT_TMA': '0'}, {'ADMP_FFT_XWIDE': '0'}])
@pytest.mark.parametrize('K', [(308, 616, 154), (616, 154, 308), (154, 1232, 616)])
def test_tile_width_and_tma_switches_give_the_same_round_trip(K, env, monkeypatch):
    """Every tile-load path of the float64 passes (TMA tensor-map copies, TMA bulk copies, cp.async; wide and narrow tiles,
    whose tensor-map boxes must start on 128-byte shared-memory boundaries) against the cuFFT + convolve_kernel round trip."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cx = _ctx(K)
    cx.set_topology(4, None, None, None)
    sp, p = _lib.stream_ptr, _lib.ptr
    box = torch.diag(torch.tensor([31.0, 29.0, 37.0], dtype=torch.float64, device='cuda'))
    pos = torch.rand((4, 3), dtype=torch.float64, device='cuda') * 20
    Q = torch.rand((4, 1), dtype=torch.float64, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(5)
    mesh = torch.randn(K, dtype=torch.float64, device='cuda', generator=g)
    out = {}
    for backend in (1, 0):
        _lib.check(cx.lib.admp_ctx_set_fft_backend(cx.handle, backend))
        _lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(Q), 1, 1, None))
        _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(mesh), mesh.numel() * 8, 1))
        scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device='cuda')
        if backend == 1:
            _lib.check(cx.lib.admp_pme_fft_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
        else:
            _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 0))
            _lib.check(cx.lib.admp_pme_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
            _lib.check(cx.lib.admp_pme_fft(cx.handle, sp(), 1))
        phi = torch.empty_like(mesh)
        _lib.check(cx.lib.admp_ctx_buffer_io(cx.handle, sp(), 0, p(phi), phi.numel() * 8, 0))
        torch.cuda.synchronize()
        out[backend] = (phi, scal.clone())
    a, b = out[1], out[0]
    assert (a[0] - b[0]).abs().max().item() < 1e-11 * b[0].abs().max().item()
    assert abs(a[1][_lib.S_E_RECIP].item() - b[1][_lib.S_E_RECIP].item()) < 1e-11 * abs(b[1][_lib.S_E_RECIP].item())


def test_in_place_mesh_equals_separate_mesh(monkeypatch):
    """The fused evaluations keep the real mesh in the spectrum buffer (line by line in place, ADMP_MESH_INPLACE, default on);
    ADMP_MESH_INPLACE=0 selects the separate mesh buffer: same energies, gradients, dipoles and SCF cycle counts (the two differ
    only in the order of the spread atomics), polarizable and dispersion, on a mesh with three different fast sizes."""
    from oracle import fixtures, pairlist
    from admp_b200.pme import ADMPPmeForce
    from admp_b200.disp_pme import ADMPDispPmeForce
    s = fixtures.lattice_water(4, 3.15, seed=5)
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 5.0)
    out = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('ADMP_MESH_INPLACE', mode)
        calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
        calc.update_env('K1', 154); calc.update_env('K2', 308); calc.update_env('K3', 154)
        E, g, vir = calc.get_forces_and_virial(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        E1 = calc.get_energy(s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
        disp = ADMPDispPmeForce(s.box, s.covalent_map, 5.0, 1e-4, 10)
        disp.update_env('K1', 154); disp.update_env('K2', 154); disp.update_env('K3', 308)
        c_list = torch.as_tensor(np.tile(np.array([[37.2, 85.3, 134.4], [7.6, 11.9, 15.1], [7.6, 11.9, 15.1]]), (s.n_atoms // 3, 1)))
        Ed, gd = disp.get_forces(s.positions, s.box, pairs, c_list, s.mScales)
        out[mode] = (E.item(), g.cpu(), vir.cpu(), calc.U_ind.cpu(), calc.n_cycle, E1.item(), Ed.item(), gd.cpu())
    a, b = out['1'], out['0']
    assert a[4] == b[4]
    for k in (0, 5, 6):
        assert abs(a[k] - b[k]) <= 1e-11 * abs(b[k])
    for k in (1, 2, 3, 7):
        assert (a[k] - b[k]).abs().max().item() <= 1e-10 * b[k].abs().max().item()


def test_in_flight_hint_changes_kernel_shapes_not_results():
    """admp_ctx_set_in_flight(n > 1) (set by parallel.sibling_calculators for frame batches) selects the narrow gather: same
    energies, gradients, dipoles and cycle counts up to summation order."""
    from oracle import fixtures, pairlist
    from admp_b200.pme import ADMPPmeForce
    s = fixtures.lattice_water(4, 3.15, seed=5)
    pairs, _ = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 5.0)
    calc = ADMPPmeForce(s.box, s.axis_type, s.axis_indices, s.covalent_map, 5.0, 1e-4, 2, lpol=True)
    args = (s.positions, s.box, pairs, s.Q_local, s.pol, s.tholes, s.mScales, s.pScales, s.dScales)
    out = []
    for n in (1, 4, 1):
        _lib.check(calc._ctx.lib.admp_ctx_set_in_flight(calc._ctx.handle, n))
        E, g, vir = calc.get_forces_and_virial(*args)
        out.append((E.item(), g.cpu(), vir.cpu(), calc.U_ind.cpu(), calc.n_cycle))
    for a in out[1:]:
        assert a[4] == out[0][4] and abs(a[0] - out[0][0]) <= 1e-11 * abs(out[0][0])
        for k in (1, 2, 3):
            assert (a[k] - out[0][k]).abs().max().item() <= 1e-10 * out[0][k].abs().max().item()
    assert calc._ctx.lib.admp_ctx_set_in_flight(calc._ctx.handle, 0) != 0      # rejected
