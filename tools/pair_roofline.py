#!/usr/bin/env python
"""Supplementary kernel rooflines on the dense (liquid-density) water box of SURVEY 8(d):
    python tools/pair_roofline.py [n_side=32] [rc=8.0]
Times the real-space multipole pair kernel (energy + all adjoints, polarizable) and its field-only SCF variant
with CUDA events and reports pairs/s, algorithmic FLOP/s (1 719 flop per polarizable pair for
energy + adjoints) against the measured FP64 FMA peak (tools/fp_peak.cu: 33.1 TFLOP/s on this pool's B200);
then spread / gather on the same box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                  # noqa: E402
import torch                                        # noqa: E402
from admp_b200 import _lib, workloads               # noqa: E402
from admp_b200._ctx import Context, to_dev          # noqa: E402
from admp_b200.neighbor import neighbor_list        # noqa: E402

FP64_PEAK = float(os.environ.get('FP64_PEAK_TFLOPS', '33.1'))
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rc = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
w = workloads.dense_water(n_side)
L = w.box[0, 0]
kappa = np.sqrt(-np.log(2e-4)) / rc
K = 154 * max(1, int(round(L / 99.3)))
cx = Context()
cx.set_topology(w.n_atoms, w.axis_type, w.axis_indices, w.covalent_map)
cx.set_pme(kappa, K, K, K, 2)
dt, dev = cx.dtype, cx.device
pos, box, Ql, pol, th, mS, pS = (to_dev(x, dt, dev) for x in (w.positions, w.box, w.Q_local, w.pol, w.tholes, w.mScales, w.pScales))
n = w.n_atoms
nb = neighbor_list(w.box, rc).allocate(w.positions)
pairs, npairs = nb.pairs, nb.n_pairs
rows = int(pairs.shape[0])
p, sp = _lib.ptr, _lib.stream_ptr
M = torch.empty((n, 10), dtype=dt, device=dev)
_lib.check(cx.lib.admp_frames_fwd(cx.handle, sp(), p(pos), p(box), p(Ql), p(M), None, None))
g = torch.Generator(device='cuda').manual_seed(1)
U = 0.01 * torch.randn((n, 3), dtype=dt, device=dev, generator=g)
scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
dpos = torch.zeros((n, 3), dtype=dt, device=dev)
G = torch.zeros((n, 10), dtype=dt, device=dev)
F = torch.zeros((n, 3), dtype=dt, device=dev)
print('dense box: %d atoms, L = %.2f A, rc = %.1f A, %d pairs (%.1f per atom), mesh %d^3' % (n, L, rc, npairs, 2.0 * npairs / n, K))


def timed(fn, reps=5):
    _lib.check(fn())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.check(fn())
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


fl = _lib.WANT_GRAD | _lib.WANT_VIRIAL | (0x40000000 if os.environ.get('PAIR_NO_ISIDE') else 0)
for force, label in ((-1, 'flat rows'), (1, 'cluster tiles')):
    _lib.check(cx.lib.admp_ctx_set_pair_cluster(cx.handle, force, 0))
    ms = timed(lambda: cx.lib.admp_pme_real(cx.handle, sp(), p(pos), p(box), p(pairs), rows, p(M), p(U), p(pol), p(th), p(mS), p(pS), 0, fl,
                                            p(dpos), p(G), p(F), None, None, p(scal)))
    tf = npairs * 1719 / (ms * 1e-3) / 1e12
    print('[%s, active %d] pair pass incl. tile build (E + adjoints, polarizable): %.3f ms, %.2f Gpairs/s, %.2f TFLOP/s algorithmic = %.1f %% of the FP64 FMA peak (%.1f)'
          % (label, cx.lib.admp_ctx_pair_cluster_active(cx.handle), ms, npairs / ms / 1e6, tf, 100 * tf / FP64_PEAK, FP64_PEAK))
    ms = timed(lambda: cx.lib.admp_pme_real(cx.handle, sp(), p(pos), p(box), p(pairs), rows, p(M), p(U), p(pol), p(th), p(mS), p(pS), 1, 0,
                                            None, None, p(F), None, None, p(scal)))
    print('[%s] pair pass incl. tile build (SCF field only): %.3f ms, %.2f Gpairs/s' % (label, ms, npairs / ms / 1e6))
_lib.check(cx.lib.admp_ctx_set_pair_cluster(cx.handle, -1 if os.environ.get('PAIR_FLAT') else 1, 0))
ms = timed(lambda: cx.lib.admp_pme_real(cx.handle, sp(), p(pos), p(box), p(pairs), rows, p(M), p(U), p(pol), p(th), p(mS), p(pS), 0, fl,
                                        p(dpos), p(G), p(F), None, None, p(scal)))
tf = npairs * 1719 / (ms * 1e-3) / 1e12
print('pair kernel (E + adjoints, polarizable): %.3f ms, %.2f Gpairs/s, %.2f TFLOP/s algorithmic = %.1f %% of the FP64 FMA peak (%.1f)'
      % (ms, npairs / ms / 1e6, tf, 100 * tf / FP64_PEAK, FP64_PEAK))
ms = timed(lambda: cx.lib.admp_pme_real(cx.handle, sp(), p(pos), p(box), p(pairs), rows, p(M), p(U), p(pol), p(th), p(mS), p(pS), 1, 0,
                                        None, None, p(F), None, None, p(scal)))
print('pair kernel (SCF field only): %.3f ms, %.2f Gpairs/s' % (ms, npairs / ms / 1e6))
ms = timed(lambda: cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
print('spread (zero-fill + scatter): %.3f ms; %.1f GB/s (mesh write + 216 x 16 B per atom)' % (ms, (8 * K ** 3 + 216 * 16 * n) / ms / 1e6))
_lib.check(cx.lib.admp_pme_fft_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
ms = timed(lambda: cx.lib.admp_pme_gather(cx.handle, sp(), p(pos), p(M), 10, 10, None, 0, _lib.WANT_GRAD, p(dpos), p(G), 10, None, p(scal)))
print('gather (full): %.3f ms; %.1f GB/s (min(mesh, 216 x 8 B per atom))' % (ms, min(8 * K ** 3, 216 * 8 * n) / ms / 1e6))
torch.cuda.synchronize()
