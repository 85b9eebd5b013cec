#!/usr/bin/env python
"""Summarise an Nsight Compute report (`ncu --set full ... -o X`) per kernel launch:
    python tools/ncu_summary.py X.ncu-rep [--md]
Prints duration, DRAM traffic, throughput percentages, occupancy, FP64/FMA pipe utilisation, shared
memory bank conflicts and the top warp-stall reasons (PC sampling)."""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'dur'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64cyc%'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
    ('smsp__inst_executed.sum', 'warp_inst'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bank_conf'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__block_size', 'block'),
    ('launch__grid_size', 'grid'),
    ('launch__shared_mem_per_block_dynamic', 'dsmem'),
    ('launch__occupancy_limit_registers', 'lim_reg'),
    ('launch__occupancy_limit_shared_mem', 'lim_smem'),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    kn = h.index('Kernel Name')
    for r in rows[2:]:
        print('### ' + r[kn][:110])
        vals = []
        for name, short in WANT:
            if name in h:
                i = h.index(name)
                vals.append('%s=%s%s' % (short, r[i], (' ' + units[i]) if units[i] and units[i] not in ('%',) else ''))
        print('  ' + '  '.join(vals))
        stalls = []
        for i, n in enumerate(h):
            if n.startswith('smsp__pcsamp_warps_issue_stalled_') and not n.endswith('_not_issued'):
                try:
                    stalls.append((float(r[i].replace(',', '')), n[len('smsp__pcsamp_warps_issue_stalled_'):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        stalls.sort(reverse=True)
        print('  stalls: ' + ', '.join('%s %.0f%%' % (n, 100 * v / tot) for v, n in stalls[:7]))


if __name__ == '__main__':
    main()
