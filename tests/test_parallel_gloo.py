"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in admp_b200/parallel.py: partitioning,
the packed all-reduce, and the decomposition algebra (frame sharding; atom-block real space +
linear mesh) evaluated with the CPU oracle under a real process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from admp_b200 import parallel


def test_partitions_cover_everything():
    for n, w in [(3072, 2), (3072, 8), (99, 4), (786432, 8)]:
        blocks = parallel.partition_atoms(n, w, 3)
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        assert all(f % 3 == 0 and c % 3 == 0 for f, c in blocks)
        assert all(blocks[k][0] + blocks[k][1] == blocks[k + 1][0] for k in range(w - 1))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 3
    rows = parallel.partition_rows(12272, 8)
    assert sum(c for _, c in rows) == 12272 and rows[-1][0] + rows[-1][1] == 12272
    assert sorted(sum((parallel.shard_frames(10, r, 4) for r in range(4)), [])) == list(range(10))
    with pytest.raises(ValueError):
        parallel.partition_atoms(10, 2, 3)


def test_slab_ownership_is_a_partition_by_fractional_x():
    rng = np.random.default_rng(3)
    box = np.diag([40.0, 50.0, 60.0])
    pos = rng.uniform(-80.0, 120.0, size=(999, 3))            # unwrapped positions are fine (admp/recip.py:324 wraps)
    for world in (1, 2, 8):
        owned = parallel.partition_atoms_by_slab(pos, box, world)
        allidx = np.concatenate(owned)
        assert len(owned) == world and sorted(allidx.tolist()) == list(range(999))
        for r, idx in enumerate(owned):
            sx = pos[idx, 0] / 40.0
            sx -= np.floor(sx)
            assert np.all(sx >= r / world - 1e-12) and np.all(sx < (r + 1) / world + 1e-12)
            assert np.all(np.diff(idx) > 0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import fixtures, pairlist
        from oracle import realspace as orc
        from oracle import reciprocal as orecip
        from oracle.frames import construct_local_frames
        from oracle.harmonics import rot_local2global
        # 1. packed all-reduce: several tensors, one collective
        a = torch.full((3, 2), float(rank + 1), dtype=torch.float64)
        b = torch.arange(5, dtype=torch.float32) * (rank + 1)
        parallel.allreduce_sum_([a, None, b])
        assert torch.all(a == 3.0) and torch.allclose(b, torch.arange(5, dtype=torch.float32) * 3)

        s = fixtures.lattice_water(3, 3.3, seed=4)
        pairs, npairs = pairlist.build_pairs(s.positions.numpy(), s.box.numpy(), 4.9)
        pairs = pairs[:npairs]
        kappa, K = 0.55, (24, 24, 24)
        # 2. frame sharding: every frame evaluated exactly once, parameter gradients summed
        frames = [s.jitter(1000 + f) for f in range(5)]
        mine = parallel.shard_frames(len(frames), rank, world)
        mS = s.mScales.clone().requires_grad_(True)
        g_local = torch.zeros(5, dtype=torch.float64)
        e_local = torch.zeros(len(frames), dtype=torch.float64)
        for f in mine:
            E = orc.energy_pme(frames[f], s.box, pairs, s.Q_local, None, None, None, mS, None, None, s.covalent_map, s.axis_type,
                               s.axis_indices, kappa, *K, 2, False)
            g_local += torch.autograd.grad(E, mS)[0]
            e_local[f] = E.detach()
        parallel.allreduce_sum_([g_local, e_local])
        # 3. atom-block algebra: pair energy over row slices + mesh linear in atoms
        fr = construct_local_frames(s.positions, s.box, s.axis_type, s.axis_indices)
        Qg = rot_local2global(s.Q_local, fr, 2)
        r0, rc = parallel.partition_rows(npairs, world)[rank]
        e_real = orc.pme_real(s.positions, s.box, pairs[r0:r0 + rc], Qg, None, None, None, s.mScales, None, None, s.covalent_map,
                              kappa, 2, False).reshape(1)
        a0, ac = parallel.partition_atoms(s.n_atoms, world, 3)[rank]
        mesh = orecip.spread(s.positions[a0:a0 + ac], s.box, Qg[a0:a0 + ac], K, 2)
        parallel.allreduce_sum_([e_real, mesh])
        # 4. x-slab algebra (SlabPme / admp_slab_*): atoms owned by fractional x, planes [floor(r K1/P), floor((r+1) K1/P)),
        #    Z and Y transforms on the own planes, the X transform on the own share of the (y, kz) columns after the
        #    exchange of planes - on an uneven split (K1 = 25 over 2 ranks: 12 + 13 planes)
        Ks = (25, 24, 24)
        owned = parallel.partition_atoms_by_slab(s.positions.numpy(), s.box.numpy(), world)[rank]
        idx = torch.as_tensor(owned)
        part = orecip.spread(s.positions[idx], s.box, Qg[idx], Ks, 2)             # own atoms, anywhere on the mesh
        parallel.allreduce_sum_([part])                                            # = what the peer atomics accumulate
        x0, x1 = Ks[0] * rank // world, Ks[0] * (rank + 1) // world
        slab_zy = torch.fft.fft(torch.fft.rfft(part[x0:x1], dim=2), dim=1)         # Z (real -> half spectrum), then Y
        counts = [Ks[0] * (q + 1) // world - Ks[0] * q // world for q in range(world)]
        pad = torch.zeros((max(counts), Ks[1], Ks[2] // 2 + 1, 2), dtype=torch.float64)      # equal-size real buffers for gloo
        pad[:x1 - x0] = torch.view_as_real(slab_zy)
        got = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(got, pad)                                                  # the pull of the fused X pass
        planes = [torch.view_as_complex(got[q][:counts[q]].contiguous()) for q in range(world)]
        cols = torch.cat(planes, dim=0).reshape(Ks[0], -1)
        ncol = cols.shape[1]
        c0, c1 = ncol * rank // world, ncol * (rank + 1) // world
        mine_x = torch.fft.fft(cols[:, c0:c1], dim=0)
        full_mesh_s = orecip.spread(s.positions, s.box, Qg, Ks, 2)
        ref_x = torch.fft.fftn(torch.fft.rfft(full_mesh_s, dim=2), dim=(0, 1)).reshape(Ks[0], -1)[:, c0:c1]
        slab_err = torch.tensor([(mine_x - ref_x).abs().max().item() / ref_x.abs().max().item()], dtype=torch.float64)
        dist.all_reduce(slab_err, op=dist.ReduceOp.MAX)
        if rank == 0:
            ref_g = torch.zeros(5, dtype=torch.float64)
            ref_e = []
            for f in range(len(frames)):
                E = orc.energy_pme(frames[f], s.box, pairs, s.Q_local, None, None, None, mS, None, None, s.covalent_map, s.axis_type,
                                   s.axis_indices, kappa, *K, 2, False)
                ref_g += torch.autograd.grad(E, mS)[0]
                ref_e.append(E.item())
            full_real = orc.pme_real(s.positions, s.box, pairs, Qg, None, None, None, s.mScales, None, None, s.covalent_map, kappa, 2, False)
            full_mesh = orecip.spread(s.positions, s.box, Qg, K, 2)
            out.put(dict(g=(g_local - ref_g).abs().max().item() / ref_g.abs().max().item(),
                         e=float(np.abs(e_local.numpy() - np.array(ref_e)).max()),
                         real=abs(e_real.item() - full_real.item()) / abs(full_real.item()),
                         mesh=(mesh - full_mesh).abs().max().item(), slab=slab_err.item()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    res = None
    for _ in range(300):                      # fail fast when a worker dies instead of waiting for the full timeout
        try:
            res = out.get(timeout=1)
            break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    assert res is not None, 'a gloo worker failed: exit codes %s' % [p.exitcode for p in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res['g'] < 1e-12 and res['e'] < 1e-9 and res['real'] < 1e-12 and res['mesh'] < 1e-12
    assert res['slab'] < 1e-12
