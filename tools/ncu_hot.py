#!/usr/bin/env python
"""Hot instructions of one kernel from an ncu report (source page, SASS view):
    python tools/ncu_hot.py X.ncu-rep kernel_regex [launch_index] [min_pct]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx, '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
print(rows[0][1][:120])
h = rows[hi]
si, sa, ie = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) <= sa or not r[0].startswith('0x'):
        continue
    data.append((int(r[sa] or 0), r[si].strip(), int(r[ie] or 0)))
tot = sum(d[0] for d in data) or 1
print('samples %d, instructions %d, executed warp-instr %d' % (tot, len(data), sum(d[2] for d in data)))
for i, (s, src, n) in enumerate(data):
    if s >= tot * minp / 100:
        print('%5d %6d %5.1f%% exec=%7d  %s' % (i, s, 100 * s / tot, n, src[:100]))
