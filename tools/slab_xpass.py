#!/usr/bin/env python
"""Times the three phases of the x-slab FFT round trip on N ranks (torchrun), device time, max over ranks:
    torchrun --nproc-per-node N tools/slab_xpass.py C5 [reps]
Environment switches read by the library: ADMP_SLAB_PULL=0 (in-kernel peer loads), ADMP_SLAB_CHUNKS=k,
ADMP_SLAB_SKIP=1|2 (diagnosis: no copies | no kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                        # noqa: E402
import torch.distributed as dist                    # noqa: E402
from admp_b200 import _lib, workloads               # noqa: E402
from admp_b200.parallel import SlabPme              # noqa: E402
from admp_b200.pme import ADMPPmeForce              # noqa: E402

REPS = {'C2': (1, 1, 1), 'C3': (2, 4, 4), 'C5': (4, 8, 8)}
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C3'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
base = workloads.water_box((1, 1, 1), polarizable=True)
K = tuple(154 * r for r in REPS[cfg])
calc = ADMPPmeForce(base.box, base.axis_type, base.axis_indices, base.covalent_map, base.rc, base.ethresh, 2, lpol=True)
calc.update_env('kappa', base.kappa)
for d in range(3):
    calc.update_env('K%d' % (d + 1), K[d])
sl = SlabPme(calc, rank, world)
sl._setup()
c, lib = calc._ctx, calc._ctx.lib
p, sp = _lib.ptr, _lib.stream_ptr
box = torch.diag(torch.tensor([50.0 * r for r in REPS[cfg]], dtype=torch.float64, device='cuda'))
_lib.check(lib.admp_set_box(c.handle, sp(), p(box)))
mesh = c.mesh_view(K)
mesh.normal_()
scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device='cuda')
tok = torch.zeros(1, device='cuda')
names = ['Z+Y forward (own planes)', 'fused X pass (peer)', 'Y+Z inverse (own planes)']
tot = [0.0, 0.0, 0.0]
for it in range(reps + 1):
    for ph in range(3):
        dist.all_reduce(tok)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.admp_slab_fft(c.handle, sp(), ph, _lib.CK_COULOMB, 0, p(scal)))
        b.record()
        b.synchronize()
        if it:
            tot[ph] += a.elapsed_time(b)
t = torch.tensor(tot, device='cuda') / reps
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    sb = 16 * K[0] * K[1] * (K[2] // 2 + 1)
    print('%s mesh %s on %d ranks, env PULL=%s CHUNKS=%s SKIP=%s' % (cfg, K, world, os.environ.get('ADMP_SLAB_PULL'), os.environ.get('ADMP_SLAB_CHUNKS'),
                                                                   os.environ.get('ADMP_SLAB_SKIP')))
    for ph in range(3):
        print('  %-28s %8.3f ms' % (names[ph], t[ph].item()))
    x = sb / world * (world - 1) / world
    print('  X pass: %.2f GB pulled + %.2f GB pushed per rank -> %.0f GB/s each way' % (x / 1e9, x / 1e9, x / t[1].item() / 1e6))
dist.barrier()
sl.close()
dist.destroy_process_group()
