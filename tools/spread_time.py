#!/usr/bin/env python
"""Spread variants timed with CUDA events (L2 flushed before every launch):
    python tools/spread_time.py c2|c3|c5|dense [iters]
per-atom scatter (+ zero-fill), brick-staged with 16- and 32-deep bricks (with the per-evaluation binning, and on reused bins)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                              # noqa: E402
import torch                                    # noqa: E402
from admp_b200 import _lib, workloads           # noqa: E402
from admp_b200._ctx import Context, to_dev      # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'c2'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if which == 'dense':
    w = workloads.dense_water(64)
    K = (308, 308, 308)
    kappa = float(np.sqrt(-np.log(2e-4)) / 8.0)
else:
    w = workloads.water_box({'c2': (1, 1, 1), 'c3': (2, 4, 4), 'c5': (4, 8, 8)}[which], polarizable=True)
    K, kappa = tuple(w.K), w.kappa
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
p, sp = _lib.ptr, _lib.stream_ptr


def timed(fn):
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(fn())
        b.record()
        b.synchronize()
        if it >= 2:
            ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


ref = None
G = K[0] * K[1] * K[2]
for label, env in (('per-atom', None), ('bricks z16', '16'), ('bricks z32', '32')):
    if env:
        os.environ['ADMP_BRICK_Z'] = env
    cx = Context()
    cx.set_topology(w.n_atoms, w.axis_type, w.axis_indices, w.covalent_map)
    cx.set_pme(kappa, K[0], K[1], K[2], 2)
    _lib.check(cx.lib.admp_ctx_set_spread(cx.handle, 1 if env else 0))
    dt, dev = cx.dtype, cx.device
    pos, box, Ql = (to_dev(x, dt, dev) for x in (w.positions, w.box, w.Q_local))
    n = w.n_atoms
    M = torch.empty((n, 10), dtype=dt, device=dev)
    _lib.check(cx.lib.admp_frames_fwd(cx.handle, sp(), p(pos), p(box), p(Ql), p(M), None, None))
    t_all = timed(lambda: cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
    t_only = timed(lambda: cx.lib.admp_pme_spread_only(cx.handle, sp(), p(pos), p(M), 10, 10, None))
    _lib.check(cx.lib.admp_pme_spread(cx.handle, sp(), p(pos), p(box), p(M), 10, 10, None))
    torch.cuda.synchronize()
    mesh = cx.mesh_view(K).clone()
    if ref is None:
        ref = mesh
    err = (mesh - ref).abs().max().item() / ref.abs().max().item()
    contract = (8 * G + 13 * 8 * n) / 1e6
    print('%-11s %s mesh %dx%dx%d %7d atoms: full %8.4f ms  kernel-only %8.4f ms  (%6.0f GB/s on w*G + 13*w*Na)  bricks=%d  max rel diff %.1e'
          % (label, which, K[0], K[1], K[2], n, t_all, t_only, contract / t_all, cx.lib.admp_ctx_spread_bricks(cx.handle), err))
    cx.close()
    del mesh
