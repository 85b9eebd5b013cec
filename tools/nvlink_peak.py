#!/usr/bin/env python
"""Measured NVLink peak of this box for the x-slab FFT's exchange pattern: every GPU pulls a slice from EVERY other GPU at
the same time (all-pairs peer copies, one stream per (destination, source) pair, DMA engines), the pattern of the
distributed transpose inside the fused X pass (admp_b200/csrc/fft.cu, slab phase 1).

    python tools/nvlink_peak.py [n_gpus=all] [MiB per pair=256]

One process, all GPUs visible. Prints per-GPU ingress GB/s (sum over sources / device time, CUDA events on the destination)
and writes gpurun_out/nvlink_peak_<N>gpu.json. Also times a single pair in one direction (the number the profiling guide
quotes as 770 GB/s per direction)."""
import json
import os
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = min(n, torch.cuda.device_count())
nbytes = mib * 1024 * 1024
src = {}
dst = {}
streams = {}
for d in range(n):
    with torch.cuda.device(d):
        src[d] = torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % d).fill_(d + 1)
        for s in range(n):
            if s != d:
                dst[(d, s)] = torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % d)
                streams[(d, s)] = torch.cuda.Stream(device=d)


def sync_all():
    for d in range(n):
        torch.cuda.synchronize(d)


def all_pairs(reps):
    start, stop = {}, {}
    sync_all()
    for d in range(n):
        with torch.cuda.device(d):
            start[d] = torch.cuda.Event(enable_timing=True)
            stop[d] = torch.cuda.Event(enable_timing=True)
            start[d].record(torch.cuda.current_stream(d))
            for s in range(n):
                if s != d:
                    streams[(d, s)].wait_event(start[d])
    for _ in range(reps):
        for d in range(n):
            for k in range(1, n):
                s = (d + k) % n                        # staggered source order
                with torch.cuda.stream(streams[(d, s)]):
                    dst[(d, s)].copy_(src[s], non_blocking=True)
    for d in range(n):
        with torch.cuda.device(d):
            cur = torch.cuda.current_stream(d)
            for s in range(n):
                if s != d:
                    e = torch.cuda.Event()
                    e.record(streams[(d, s)])
                    cur.wait_event(e)
            stop[d].record(cur)
    sync_all()
    return [start[d].elapsed_time(stop[d]) for d in range(n)]


out = dict(n_gpus=n, mib_per_pair=mib)
if n >= 2:
    all_pairs(2)
    reps = 10
    ms = all_pairs(reps)
    ingress = [reps * (n - 1) * nbytes / (t * 1e-3) / 1e9 for t in ms]
    out['all_pairs_ingress_gbs_per_gpu'] = [round(x, 1) for x in ingress]
    out['all_pairs_ingress_gbs_min'] = round(min(ingress), 1)
    # one pair, one direction
    with torch.cuda.device(0):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst[(0, 1)].copy_(src[1])
        torch.cuda.synchronize(0)
        a.record()
        for _ in range(reps):
            dst[(0, 1)].copy_(src[1], non_blocking=True)
        b.record()
        b.synchronize()
        out['single_pair_one_direction_gbs'] = round(reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9, 1)
    assert int(dst[(0, 1)][0].item()) == 2
print(json.dumps(out))
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/nvlink_peak_%dgpu.json' % n, 'w') as f:
    f.write(json.dumps(out) + '\n')
