#!/usr/bin/env python
"""Small driver for profiling the reciprocal-space stages on a water_box(reps) mesh:
    python tools/prof_recip.py [nx ny nz] [iters]
Runs spread -> fused FFT/convolve round trip -> gather `iters` times (and times them with CUDA events)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                    # noqa: E402
from admp_b200 import _lib, workloads           # noqa: E402
from admp_b200._ctx import Context, to_dev      # noqa: E402

reps = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 1, 1)
iters = int(sys.argv[4]) if len(sys.argv) >= 5 else 5
w = workloads.water_box(reps, polarizable=True)
cx = Context()
cx.set_topology(w.n_atoms, w.axis_type, w.axis_indices, w.covalent_map)
cx.set_pme(w.kappa, w.K[0], w.K[1], w.K[2], 2)
dt, dev = cx.dtype, cx.device
pos, box, Ql = (to_dev(x, dt, dev) for x in (w.positions, w.box, w.Q_local))
n = w.n_atoms
M = torch.empty((n, 10), dtype=dt, device=dev)
p, sp = _lib.ptr, _lib.stream_ptr
_lib.check(cx.lib.admp_frames_fwd(cx.handle, sp(), p(pos), p(box), p(Ql), p(M), None, None))
scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
dpos = torch.zeros((n, 3), dtype=dt, device=dev)
G = torch.zeros((n, 10), dtype=dt, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(name, fn):
    ts = []
    for it in range(iters + 1):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(fn())
        b.record()
        b.synchronize()
        if it:
            ts.append(a.elapsed_time(b))
    print('%-28s %9.4f ms (mesh %dx%dx%d, backend %d)' % (name, sum(ts) / len(ts), w.K[0], w.K[1], w.K[2],
                                                           cx.lib.admp_ctx_fft_backend(cx.handle)))


def warm(name, fn, reps=20):
    """back-to-back launches, no L2 flush in between (the in-loop regime of the SCF cycle)"""
    _lib.check(fn())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.check(fn())
    b.record()
    b.synchronize()
    print('%-28s %9.4f ms warm (back to back x%d)' % (name, a.elapsed_time(b) / reps, reps))


if os.environ.get('PROF_ONCE'):
    # one launch of every reciprocal-space kernel (for `ncu --set full`): spread, the five FFT passes, gather
    _lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
    _lib.check(cx.lib.admp_pme_fft_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
    _lib.check(cx.lib.admp_pme_gather(cx.handle, sp(), p(pos), p(M), 10, 10, None, 0, _lib.WANT_GRAD, p(dpos), p(G), 10, None, p(scal)))
    torch.cuda.synchronize()
    print('one launch per kernel on mesh %dx%dx%d' % tuple(w.K))
    sys.exit(0)

if os.environ.get('PROF_WARM'):
    warm('spread (zero + scatter)', lambda: cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
    warm('fused roundtrip', lambda: cx.lib.admp_pme_fft_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
    for k, nm in enumerate(['pass z_fwd', 'pass y_fwd', 'pass x_conv', 'pass y_inv', 'pass z_inv']):
        warm(nm, lambda k=k: cx.lib.admp_pme_fft_pass(cx.handle, sp(), k, _lib.CK_COULOMB, p(scal)))
    Fs = torch.zeros((n, 3), dtype=dt, device=dev)
    warm('gather (field only)', lambda: cx.lib.admp_pme_gather(cx.handle, sp(), p(pos), p(M), 10, 10, None, 1, 0, None, None, 10, p(Fs), p(scal)))
    warm('gather (full)', lambda: cx.lib.admp_pme_gather(cx.handle, sp(), p(pos), p(M), 10, 10, None, 0, _lib.WANT_GRAD, p(dpos), p(G), 10,
                                                       None, p(scal)))
    torch.cuda.synchronize()
    sys.exit(0)

timed('spread (zero + scatter)', lambda: cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
if cx.lib.admp_ctx_fft_backend(cx.handle):
    timed('fused fft+convolve roundtrip', lambda: cx.lib.admp_pme_fft_convolve(cx.handle, sp(), _lib.CK_COULOMB, 0, p(scal)))
if cx.lib.admp_ctx_fft_backend(cx.handle):
    for k, nm in enumerate(['pass z_fwd', 'pass y_fwd', 'pass x_conv', 'pass y_inv', 'pass z_inv']):
        timed(nm, lambda k=k: cx.lib.admp_pme_fft_pass(cx.handle, sp(), k, _lib.CK_COULOMB, p(scal)))
timed('fft forward', lambda: cx.lib.admp_pme_fft(cx.handle, sp(), 0))
timed('fft inverse', lambda: cx.lib.admp_pme_fft(cx.handle, sp(), 1))
timed('gather (full)', lambda: cx.lib.admp_pme_gather(cx.handle, sp(), p(pos), p(M), 10, 10, None, 0, _lib.WANT_GRAD, p(dpos), p(G), 10,
                                                      None, p(scal)))
torch.cuda.synchronize()
