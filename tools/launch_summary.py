#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`):
    python tools/launch_summary.py X.csv [title] > profiles/rNN_launch_summary.md
Times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES, not absolute values."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[mv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[mu], 1.0)
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if 'admp::' in k)
    print('# %s' % title)
    print()
    print('%d launches, %.1f us total (cold-cache, serialised: compare shares); hand-written admp:: kernels = %.1f %% of the time.' % (
        n, tot, 100 * ours / tot))
    print()
    print('| launches | total us | share | avg us | kernel |')
    print('|---:|---:|---:|---:|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %d | %.1f | %.1f%% | %.2f | `%s` |' % (a[0], a[1], 100 * a[1] / tot, a[1] / a[0], k[:110]))


if __name__ == '__main__':
    main()
