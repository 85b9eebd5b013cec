#!/usr/bin/env python
"""SASS opcode counts per kernel of admp_b200/lib/libadmp_b200.so (cuobjdump -sass; no GPU needed):
    python tools/sass_opcodes.py [regex of demangled kernel names] > profiles/sass_opcodes_<round>.md
Default: the float64 kernels the default path launches (the names of the committed ncu launch list) + the opt-in brick spread."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'admp_b200', 'lib', 'libadmp_b200.so')
OPS = ['UTMALDG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'DFMA', 'DMUL', 'DADD', 'FFMA', 'RED', 'ATOMG', 'ATOMS', 'SHFL', 'BAR', 'LDS', 'STS', 'LDG', 'STG',
       'LDL', 'STL', 'LDC', 'LDCU', 'UMOV', 'MUFU']
DEFAULT = (r'fast_x_conv_kernel<double, 11, (14, 1, 8, 14|7, 4, 8, 28|7, 8, 8, 56), true, false, true|'
           r'fast_strided_kernel<double, 11, (14, 1|7, 8|7, 16), (1|-1), (8, 14|4, 32|4, 112), true>|'
           r'fast_z_(fwd|inv)_kernel<double, 11, 7, (1, 16, 7|4, 4, 28|8, 2, 56)>|'
           r'spread_kernel<double, true, false>|gather_kernel<double, true, [01], false, (4|16)>|spread_brick_kernel<double, true, 32>|'
           r'brick_prep_kernel<double, true>|pme_cluster_kernel<double, true, [01], (true|false)>|pme_pair_kernel<double, true, [01]>|'
           r'nb_pairs_warp_kernel<double|scf_|self_kernel<double|frames_(fwd|bwd)_kernel<double|disp_pair_kernel<double|tt_pair_kernel<double')


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else DEFAULT)
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    names = {}
    cur = None
    counts = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur:
            counts[cur]['_n'] += 1
            counts[cur][m.group(1)] += 1
    dem = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(counts, dem):
        names[k] = re.sub(r'\(.*$', '', d).replace('admp::', '').replace('(bool)1', 'true').replace('(bool)0', 'false').replace('(int)', '')
    print('# SASS opcode counts per kernel of admp_b200/lib/libadmp_b200.so (`cuobjdump -sass`, sm_100a; `tools/sass_opcodes.py`)\n')
    print('UTMALDG = `cp.async.bulk.tensor` (TMA tensor-map tile load), UBLKCP = `cp.async.bulk.shared.global` (TMA 1-D bulk copy), SYNCS = mbarrier '
          'operations, LDGSTS = `cp.async`; DFMA / DMUL / DADD = FP64 pipe; RED / ATOMG = global reductions; LDC / LDCU = constant-bank loads '
          '(butterfly constants), UMOV = immediates through the uniform datapath. No tensor-core opcodes by design (FP64 butterflies and pair '
          'arithmetic are not dense contractions). Listed: the float64 kernels the default path launches and the opt-in brick spread.\n')
    print('| kernel | instrs | ' + ' | '.join(OPS) + ' |')
    print('|---|---:|' + '---:|' * len(OPS))
    for k, c in counts.items():
        n = names[k]
        if not pat.search(n):
            continue
        print('| `%s` | %d | %s |' % (n[:110], c['_n'], ' | '.join(str(c[o]) for o in OPS)))


if __name__ == '__main__':
    main()
