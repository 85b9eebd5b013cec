#!/usr/bin/env python
"""Config C4 on one GPU: throughput of a frame batch against the number of frames in flight.
    python tools/frames_in_flight.py [n_frames=48] [max_lanes=8] [repeats=2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
from admp_b200 import workloads                     # noqa: E402
from admp_b200.neighbor import neighbor_list        # noqa: E402
from admp_b200.parallel import evaluate_frames      # noqa: E402
from admp_b200.pme import ADMPPmeForce              # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 48
mx = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = workloads.water_box((1, 1, 1), polarizable=True)
calc = ADMPPmeForce(w.box, w.axis_type, w.axis_indices, w.covalent_map, w.rc, w.ethresh, 2, lpol=True)
calc.update_env('kappa', w.kappa)
frames = [workloads.jitter_frame(w, 5000 + f) for f in range(nb)]
nl = neighbor_list(w.box, w.rc)
prs = [nl.allocate(f).pairs for f in frames]
args = (w.box, lambda f: prs[f], w.Q_local, w.pol, w.tholes, w.mScales, w.pScales)
for lanes in range(1, mx + 1):
    evaluate_frames(calc, frames[:2 * lanes], *args, in_flight=lanes)
    torch.cuda.synchronize()
    out = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        evaluate_frames(calc, frames, *args, in_flight=lanes)
        b.record()
        b.synchronize()
        out.append(nb / (a.elapsed_time(b) * 1e-3))
    graphs = [c._ctx.scf_graph_active for c in [calc] + getattr(calc, '_siblings', [])[:lanes - 1]]
    print('%d frames in flight: %s evals/s   (SCF graphs active: %s)' % (lanes, ', '.join('%.1f' % x for x in out), graphs))
