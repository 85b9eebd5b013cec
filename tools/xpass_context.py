#!/usr/bin/env python
"""How long the fused X pass takes in context: alone back to back, behind a Y-forward pass on a freshly transformed spectrum, and
as the difference (Y-forward + X) - Y-forward, with the SM clocks sampled during each series:  python tools/xpass_context.py 2 4 4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
import bench                                        # noqa: E402
from admp_b200 import _lib, workloads               # noqa: E402
from admp_b200._ctx import Context, to_dev          # noqa: E402

reps = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 4, 4)
w = workloads.water_box(reps, polarizable=True)
cx = Context()
cx.set_topology(w.n_atoms, w.axis_type, w.axis_indices, w.covalent_map)
cx.set_pme(w.kappa, w.K[0], w.K[1], w.K[2], 2)
dt, dev = cx.dtype, cx.device
pos, box, Ql = (to_dev(x, dt, dev) for x in (w.positions, w.box, w.Q_local))
M = torch.empty((w.n_atoms, 10), dtype=dt, device=dev)
p, sp = _lib.ptr, _lib.stream_ptr
_lib.check(cx.lib.admp_frames_fwd(cx.handle, sp(), p(pos), p(box), p(Ql), p(M), None, None))
scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
P = lambda k: _lib.check(cx.lib.admp_pme_fft_pass(cx.handle, sp(), k, _lib.CK_COULOMB, p(scal)))
fresh = lambda: (_lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None)), P(0))


def series(name, body, n=20, bracket=None):
    smp = bench.ClockSampler(0)
    body()
    torch.cuda.synchronize()
    smp.mark_start()
    tot = 0.0
    for _ in range(n):
        if bracket:
            bracket()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        body()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    smp.mark_stop()
    c = smp.stop()
    print('%-44s %8.4f ms   sm %s MHz %s' % (name, tot / n, c.get('sm_mhz'), c.get('reasons')))
    return tot / n


x = series('X alone, same buffer again and again', lambda: P(2))
y = series('Y-forward alone (fresh Z-forward output)', lambda: P(1), bracket=fresh)
yx = series('Y-forward + X (fresh Z-forward output)', lambda: (P(1), P(2)), bracket=fresh)
print('X in context = (Y + X) - Y = %.4f ms; alone %.4f ms' % (yx - y, x))
