#!/usr/bin/env python
"""Times the five FFT passes (and spread / gather) on the mesh of water_box(reps):  python tools/xpass_time.py 2 4 4
(the table bench.py prints as kernels.C3, for A/B runs under different ADMP_FFT_* settings)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
import bench                                        # noqa: E402
from admp_b200 import _lib                          # noqa: E402

reps = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 4, 4)
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
peak, _ = bench.measured_peaks()
smp = bench.ClockSampler(0)
smp.mark_start()
tab = bench.kernel_rooflines(torch, _lib, reps, peak, lambda: flush_buf.zero_(), n_launch=5)
smp.mark_stop()
print('clocks during the timed launches:', smp.stop())
print('settings:', {k: v for k, v in os.environ.items() if k.startswith('ADMP_')})
for k, v in tab.items():
    print('%-55s %8.4f ms  frac %s' % (k, v['ms'], v.get('frac')))
