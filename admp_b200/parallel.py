"""Multi-GPU partitioning of the PME path; one process per GPU, torch.distributed
(NCCL over NVLink on the box, gloo in the CPU tests) for the plumbing.

Two shardings, both prescribed by the north star; neither exists in the reference:

* **Frames** (config C4, force-field fitting): frames are independent units. Rank r evaluates frames
  r, r+P, ...; energies and dE/dpositions stay local; the parameter-gradient vector (dE/dQ_local,
  dE/dmScales, dE/dpScales, dE/dtholes, dE/dpol) is summed over the local frames and all-reduced once
  at the end (<= ~250 KB: latency-bound). No data-path collective.
* **Atom blocks** (config C5, >= 100k atoms): positions / multipoles are replicated; rank r owns a
  contiguous block of atoms (whole molecules) and a contiguous slice of the pair rows. Per reciprocal
  round trip: spread own atoms -> all-reduce the real mesh -> (replicated) FFT/convolution ->
  gather own atoms; pair kernel on own rows; per-atom results (field, dE/dM, dE/dr) are all-reduced.
  The replicated mesh and its all-reduce are the scaling limiter; reported as measured.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import settings


# ----------------------------------------------------------------------------------- partitioning
def shard_frames(n_frames, rank, world):
    """Round-robin frame ownership."""
    return list(range(rank, n_frames, world))


def partition_atoms(n_atoms, world, granule=3):
    """Contiguous atom blocks aligned to `granule` atoms (whole molecules): list of (first, count)."""
    units = n_atoms // granule
    if units * granule != n_atoms:
        raise ValueError('n_atoms is not a multiple of the molecule size')
    base, extra = divmod(units, world)
    out, first = [], 0
    for r in range(world):
        cnt = (base + (1 if r < extra else 0)) * granule
        out.append((first, cnt))
        first += cnt
    return out


def partition_rows(n_rows, world):
    """Contiguous, near-equal slices of the pair rows: list of (first, count)."""
    base, extra = divmod(n_rows, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def allreduce_sum_(tensors, group=None):
    """Sum a list of tensors over the ranks with ONE collective (packed flat buffer), in place."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
        off += n
    return tensors


# ----------------------------------------------------------------------------------- frames (C4)
def sibling_calculators(calc, count):
    """`count` calculators with the environment of `calc` (own context, mesh, SCF graph each), cached on `calc`:
    the lanes of the frames-in-flight batch evaluation. Index 0 is `calc` itself."""
    from .pme import ADMPPmeForce
    sibs = getattr(calc, '_siblings', None)
    key = (calc.kappa, calc.K1, calc.K2, calc.K3, calc.lmax, calc.lpol)
    if sibs is None or getattr(calc, '_siblings_key', None) != key:
        sibs = []
    while len(sibs) < count - 1:
        small = np.eye(3) * max(4.0 * float(calc.rc), 10.0)          # placeholder cell: the real K / kappa are copied below
        c2 = ADMPPmeForce(small, calc.axis_type, calc.axis_indices, calc.covalent_map, calc.rc, calc.ethresh, calc.lmax, calc.lpol)
        c2.kappa, c2.K1, c2.K2, c2.K3 = calc.kappa, calc.K1, calc.K2, calc.K3
        c2.refresh_calculators()
        sibs.append(c2)
    calc._siblings, calc._siblings_key = sibs, key
    out = [calc] + sibs[:count - 1]
    for c in out:                       # hint for the kernels: `count` evaluations share the GPU
        _lib.check(c._ctx.lib.admp_ctx_set_in_flight(c._ctx.handle, max(1, int(count))))
    return out


def evaluate_frames(calc, frames, box, pairs, Q_local, pol=None, tholes=None, mScales=None, pScales=None,
                    rank=0, world=1, group=None, param_grads=True, in_flight=1):
    """Frame-sharded evaluation with `calc` (an ADMPPmeForce). `frames`: sequence of (Na,3) arrays,
    `pairs`: one pair list or a callable frame_index -> pairs. Returns a dict with the local frame
    indices, their energies and dE/dpositions, and the frame-summed, rank-reduced parameter gradients.

    in_flight > 1 keeps that many frames in flight on one GPU (one calculator context + CUDA stream per lane):
    a 1024-water evaluation is latency-bound (a 30-cycle SCF of ~8 small dependent kernels per cycle leaves most
    of the 148 SMs idle), so independent frames overlap: measured 312 / 375 / 420 evals/s with 1 / 2 / 4 lanes (saturated: every
    FFT pass of one frame already occupies all resident-block slots). Run with CUDA_DEVICE_MAX_CONNECTIONS=32 (set before CUDA
    initialises): each lane uses three streams and with the default 8 hardware queues some lane counts serialise."""
    mine = shard_frames(len(frames), rank, world)
    dev, dt = calc._ctx.device, calc._dtype
    prep = calc._prep
    box_d, Ql = prep(box), prep(Q_local)
    polz = calc.lpol
    rest = [prep(x) for x in ((pol, tholes, mScales, pScales) if polz else (mScales,))]
    flags = _lib.WANT_GRAD | (_lib.WANT_PGRAD if param_grads else 0)
    n, nh = calc.n_atoms, (calc.lmax + 1) ** 2
    lanes = max(1, min(int(in_flight), len(mine))) if mine else 1
    calcs = sibling_calculators(calc, lanes)
    main = torch.cuda.current_stream()
    streams = [main] if lanes == 1 else [torch.cuda.Stream(device=dev) for _ in range(lanes)]

    def new_acc():
        a = dict(dQ_local=torch.zeros((n, nh), dtype=torch.float64, device=dev),
                 dmScales=torch.zeros(5, dtype=torch.float64, device=dev))
        if polz:
            a.update(dpScales=torch.zeros(5, dtype=torch.float64, device=dev),
                     dtholes=torch.zeros(n, dtype=torch.float64, device=dev),
                     dpol=torch.zeros(n, dtype=torch.float64, device=dev))
        return a

    accs = [new_acc() for _ in range(lanes)]
    zeroU = torch.zeros((n, 3), dtype=dt, device=dev) if polz else None
    from ._ctx import pairs_to_dev
    fixed_pairs = None if callable(pairs) else pairs_to_dev(pairs, dev)
    # inputs are staged on the main stream, lanes start after it
    pos_d = {f: prep(frames[f]) for f in mine}
    pr_d = {f: (pairs_to_dev(pairs(f), dev) if callable(pairs) else fixed_pairs) for f in mine}
    if lanes > 1:
        ready = torch.cuda.Event()
        ready.record(main)
        for s in streams:
            s.wait_event(ready)
    energies, grads = {}, {}
    for k, f in enumerate(mine):
        lane = k % lanes
        c = calcs[lane]
        with torch.cuda.stream(streams[lane]):
            if polz:
                r = c._eval(pos_d[f], box_d, pr_d[f], Ql, zeroU, rest[0], rest[1], rest[2], rest[3], flags, True, cache_scf=False)
            else:
                r = c._eval(pos_d[f], box_d, pr_d[f], Ql, None, None, None, rest[0], None, flags, False)
            energies[f] = r.energy
            grads[f] = r.dpos
            if param_grads:
                acc = accs[lane]
                acc['dQ_local'] += r.dQ
                acc['dmScales'] += r.scalars[_lib.S_DMSCALE:_lib.S_DMSCALE + 5]
                if polz:
                    acc['dpScales'] += r.scalars[_lib.S_DPSCALE:_lib.S_DPSCALE + 5]
                    acc['dtholes'] += r.dtholes
                    acc['dpol'] += r.dpol
    if lanes > 1:
        for s in streams:
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)
    acc = accs[0]
    for other in accs[1:]:
        for k2 in acc:
            acc[k2] += other[k2]
    if param_grads and world > 1:
        allreduce_sum_(list(acc.values()), group)
    return dict(frames=mine, energies=torch.stack([energies[f] for f in mine]) if mine else torch.zeros(0, dtype=torch.float64, device=dev),
                dpos=[grads[f] for f in mine], param_grads=acc if param_grads else None)


# ----------------------------------------------------------------------------------- atom blocks (C5)
class AtomBlockPme:
    """Atom-block decomposition of one polarizable / non-polarizable PME evaluation over the ranks of
    `group`. Every rank passes the same full inputs and receives the same full outputs."""

    def __init__(self, calc, rank=0, world=1, group=None, granule=3, emulate_blocks=None):
        """emulate_blocks=P (with world == 1) walks all P blocks in this one process, accumulating into
        the same arrays: the single-GPU check of the decomposition algebra."""
        self.calc, self.rank, self.world, self.group = calc, rank, world, group
        self.nblocks = world if emulate_blocks is None else int(emulate_blocks)
        if emulate_blocks is not None and world != 1:
            raise ValueError('emulate_blocks needs world == 1')
        self.atoms = partition_atoms(calc.n_atoms, self.nblocks, granule)
        self.mine = [rank] if emulate_blocks is None else list(range(self.nblocks))
        self._mesh = None

    # the mesh all-reduce runs in place on a zero-copy torch view of the context's mesh
    def _allreduce_mesh(self):
        import torch.distributed as dist
        if self.world == 1:
            return
        c = self.calc._ctx
        K = (self.calc.K1, self.calc.K2, self.calc.K3)
        if self._mesh is None or tuple(self._mesh.shape) != K:
            self._mesh = c.mesh_view(K)
        dist.all_reduce(self._mesh, op=dist.ReduceOp.SUM, group=self.group)

    def evaluate(self, positions, box, pairs, Q_local, pol=None, tholes=None, mScales=None, pScales=None, U_init=None,
                 want_virial=True, maxiter=None, thresh=None):
        """Returns dict(E, dpos, dbox, dQ_local, U, F, n_cycle, converged, parts)."""
        from ._ctx import pairs_to_dev
        calc = self.calc
        c, lib = calc._ctx, calc._ctx.lib
        p, sp = _lib.ptr, _lib.stream_ptr
        dev, dt = c.device, c.dtype
        n, nh = calc.n_atoms, (calc.lmax + 1) ** 2
        polz = pol is not None
        prep = calc._prep
        pos, box, Ql, mS = prep(positions).detach(), prep(box).detach(), prep(Q_local).detach(), prep(mScales).detach()
        pr = pairs_to_dev(pairs, dev)
        rows = partition_rows(int(pr.shape[0]), self.nblocks)
        blocks = [(self.atoms[b][0], self.atoms[b][1], pr[rows[b][0]:rows[b][0] + rows[b][1]].contiguous(), rows[b][1])
                  for b in self.mine]
        maxiter = settings.MAX_N_POL if maxiter is None else maxiter
        thresh = settings.POL_CONV if thresh is None else thresh
        fl = _lib.WANT_GRAD | (_lib.WANT_VIRIAL if want_virial else 0)
        vir = _lib.WANT_VIRIAL if want_virial else 0

        _lib.check(lib.admp_set_box(c.handle, sp(), p(box)))
        M = torch.empty((n, 10), dtype=dt, device=dev)
        _lib.check(lib.admp_frames_fwd(c.handle, sp(), p(pos), p(box), p(Ql), p(M), None, None))
        scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
        U = F = None
        n_cycle, conv = 0, True
        if polz:
            polt, th, pS = prep(pol).detach(), prep(tholes).detach(), prep(pScales).detach()
            U = torch.zeros((n, 3), dtype=dt, device=dev) if U_init is None else prep(U_init).detach().clone()
            F = torch.zeros((n, 3), dtype=dt, device=dev)
            state = torch.zeros(8, dtype=torch.int32, device=dev)
            for _ in range(maxiter + 1):
                scal.zero_()
                self._recip(pos, M, U, blocks, scal, 0)          # SCF cycles: energy-only (quick) X pass
                F.zero_()
                for a0, ac, my_pairs, rc in blocks:
                    _lib.check(lib.admp_pme_gather_range(c.handle, sp(), p(pos), p(M), 10, 10, p(U), 1, 0, None, None, 10, p(F), p(scal),
                                                         a0, ac))
                for a0, ac, my_pairs, rc in blocks:
                    if rc > 0:
                        _lib.check(lib.admp_pme_real(c.handle, sp(), p(pos), p(box), p(my_pairs), rc, p(M), p(U), p(polt), p(th), p(mS),
                                                     p(pS), 1, 0, None, None, p(F), None, None, p(scal)))
                allreduce_sum_([F], self.group)
                _lib.check(lib.admp_scf_step(c.handle, sp(), p(M), p(U), p(polt), p(F), int(maxiter), float(thresh), vir, p(state),
                                             p(scal)))
                st = state.cpu()
                if not int(st[5]):
                    n_cycle, conv = int(st[3]), bool(st[4])
                    break
            if want_virial:
                # final reciprocal pass on the final U with the k-space virial sums (also the refresh pass
                # after the last allowed update), as in admp_pme_eval
                scal.zero_()
                self._recip(pos, M, U, blocks, scal, vir)
        else:
            self._recip(pos, M, None, blocks, scal, vir)
        # final evaluation at fixed U; phi of the last round trip is in the context's mesh
        e_recip = scal[_lib.S_E_RECIP].clone()
        tk = scal[_lib.S_TK:_lib.S_TK + 6].clone()
        scal.zero_()
        G = torch.zeros((n, 10), dtype=dt, device=dev)
        dpos = torch.zeros((n, 3), dtype=dt, device=dev)
        dQ = torch.zeros((n, nh), dtype=dt, device=dev)
        Fo = torch.zeros((n, 3), dtype=dt, device=dev) if polz else None
        for a0, ac, my_pairs, rc in blocks:
            _lib.check(lib.admp_pme_gather_range(c.handle, sp(), p(pos), p(M), 10, 10, p(U), 0, fl, p(dpos), p(G), 10, p(Fo), p(scal),
                                                 a0, ac))
            if rc > 0:
                _lib.check(lib.admp_pme_real(c.handle, sp(), p(pos), p(box), p(my_pairs), rc, p(M), p(U), p(polt) if polz else None,
                                             p(th) if polz else None, p(mS), p(pS) if polz else None, 0, fl, p(dpos), p(G), p(Fo), None,
                                             None, p(scal)))
            _lib.check(lib.admp_pme_self_range(c.handle, sp(), p(M), p(U), p(polt) if polz else None, fl, p(G), p(Fo), None, p(scal),
                                               a0, ac))
        allreduce_sum_([G], self.group)
        for a0, ac, my_pairs, rc in blocks:
            _lib.check(lib.admp_frames_bwd_range(c.handle, sp(), p(pos), p(Ql), p(G), p(dQ), p(dpos), p(scal), a0, ac))
        allreduce_sum_([dpos, dQ, Fo, scal], self.group)
        scal[_lib.S_E_RECIP] = e_recip                     # replicated quantities: not summed
        scal[_lib.S_TK:_lib.S_TK + 6] = tk
        if want_virial:
            _lib.check(lib.admp_virial_finalize(c.handle, sp(), p(scal)))
        E = scal[_lib.S_E_REAL] + scal[_lib.S_E_RECIP] + scal[_lib.S_E_SELF] + scal[_lib.S_E_PEN]
        return dict(E=E, dpos=dpos, dbox=scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3).clone(), dQ_local=dQ, U=U, F=Fo,
                    n_cycle=n_cycle, converged=conv, scalars=scal)

    def _recip(self, pos, M, U, blocks, scal, vir):
        c, lib = self.calc._ctx, self.calc._ctx.lib
        p, sp = _lib.ptr, _lib.stream_ptr
        _lib.check(lib.admp_mesh_zero(c.handle, sp()))
        for a0, ac, _, _ in blocks:
            _lib.check(lib.admp_pme_spread_range(c.handle, sp(), p(pos), p(M), 10, 10, p(U), a0, ac))
        self._allreduce_mesh()
        if lib.admp_ctx_fft_backend(c.handle):
            _lib.check(lib.admp_pme_fft_convolve(c.handle, sp(), _lib.CK_COULOMB, vir, p(scal)))
        else:
            _lib.check(lib.admp_pme_fft(c.handle, sp(), 0))
            _lib.check(lib.admp_pme_convolve(c.handle, sp(), _lib.CK_COULOMB, vir, p(scal)))
            _lib.check(lib.admp_pme_fft(c.handle, sp(), 1))


# ----------------------------------------------------------------------------------- x-slab reciprocal space (C5)
def partition_atoms_by_slab(positions, box, world):
    """Spatial ownership for the x-slab scheme: atom a belongs to the rank whose x planes contain its fractional
    x coordinate (so nearly all of its 6x6x6 stencil is local). Returns a list of sorted int64 index arrays."""
    pos = np.asarray(positions, dtype=np.float64)
    inv = np.linalg.inv(np.asarray(box, dtype=np.float64))
    sx = (pos @ inv)[:, 0]
    sx = sx - np.floor(sx)
    owner = np.minimum((sx * world).astype(np.int64), world - 1)
    return [np.nonzero(owner == r)[0] for r in range(world)]


class SlabPme:
    """Polarizable / non-polarizable PME evaluation with reciprocal space decomposed into x slabs over the GPUs
    of one NVLink domain (include/admp_b200.h, "x-slab decomposition"): no replicated FFT, no mesh all-reduce.

    Rank r owns K1/P x planes of the mesh and of the half spectrum, the atoms whose fractional x lies in its slab
    (spread / gather), a contiguous slice of the pair rows and a contiguous atom block for the per-site stages.
    Positions, multipoles and induced dipoles stay replicated (<= 100 MB at C5); per cycle the only collectives
    are the all-reduce of the field (Na x 3) and five stream-ordered barriers; the spectrum transposes of the
    distributed FFT happen inside the fused X-pass kernel over peer-mapped memory.

    emulate_ranks=P (world == 1) walks all P ranks in one process on one device, each with its own context
    (mesh + spectrum), stage by stage: the single-GPU check of the decomposition.
    """

    def __init__(self, calc, rank=0, world=1, group=None, granule=3, emulate_ranks=None):
        self.calc, self.rank, self.world, self.group = calc, rank, world, group
        if emulate_ranks is not None and world != 1:
            raise ValueError('emulate_ranks needs world == 1')
        self.P = world if emulate_ranks is None else int(emulate_ranks)
        self.mine = [rank] if emulate_ranks is None else list(range(self.P))
        self.atoms = partition_atoms(calc.n_atoms, self.P, granule)
        self._ctxs = {}
        self._opened = []
        self._key = None
        self._tok = None
        self.profile = False
        self._spans = []

    # ---- contexts and peer tables
    def _own_buffers(self, c):
        lib = c.lib
        lib.admp_ctx_buffer.restype = ctypes.c_void_p
        return int(lib.admp_ctx_buffer(c.handle, 0)), int(lib.admp_ctx_buffer(c.handle, 1))

    def _setup(self):
        from ._ctx import Context
        calc = self.calc
        main = calc._ctx
        key = (self._own_buffers(main), calc.K1, calc.K2, calc.K3, calc.kappa)
        if key == self._key:
            return
        self.close()
        self._key = key
        lib = main.lib
        if self.world > 1:
            import torch.distributed as dist
            self._ctxs = {self.rank: main}
            hm, hs = (ctypes.create_string_buffer(64) for _ in range(2))
            mp, spp = self._own_buffers(main)
            _lib.check(lib.admp_ipc_export(ctypes.c_void_p(mp), hm))
            _lib.check(lib.admp_ipc_export(ctypes.c_void_p(spp), hs))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (hm.raw, hs.raw), group=self.group)
            mesh_ptrs, spec_ptrs = [], []
            for r, (a, b) in enumerate(gathered):
                if r == self.rank:
                    mesh_ptrs.append(mp)
                    spec_ptrs.append(spp)
                    continue
                out = []
                for h in (a, b):
                    pv = ctypes.c_void_p()
                    _lib.check(lib.admp_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(pv)))
                    self._opened.append(pv.value)
                    out.append(pv.value)
                mesh_ptrs.append(out[0])
                spec_ptrs.append(out[1])
            self._set_peers(main, self.rank, mesh_ptrs, spec_ptrs)
            self._tok = torch.zeros(1, dtype=torch.float32, device=main.device)
        else:
            self._ctxs = {0: main}
            for r in range(1, self.P):
                c = Context()
                c.set_topology(calc.n_atoms, calc.axis_type, calc.axis_indices, calc.covalent_map)
                c.set_pme(calc.kappa, calc.K1, calc.K2, calc.K3, calc.lmax)
                self._ctxs[r] = c
            bufs = [self._own_buffers(self._ctxs[r]) for r in range(self.P)]
            for r in range(self.P):
                self._set_peers(self._ctxs[r], r, [b[0] for b in bufs], [b[1] for b in bufs])

    def _set_peers(self, c, rank, mesh_ptrs, spec_ptrs):
        arr = ctypes.c_void_p * self.P
        _lib.check(c.lib.admp_ctx_set_peers(c.handle, rank, self.P, arr(*mesh_ptrs), arr(*spec_ptrs)))

    def close(self):
        lib = self.calc._ctx.lib
        for c in self._ctxs.values():
            try:
                lib.admp_ctx_set_peers(c.handle, 0, 0, None, None)
            except Exception:                                        # noqa: BLE001
                pass
        for pv in self._opened:
            lib.admp_ipc_close(ctypes.c_void_p(pv))
        self._opened = []
        for r, c in list(self._ctxs.items()):
            if c is not self.calc._ctx:
                c.close()
        self._ctxs = {}
        self._key = None

    def _barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            with self._timed('barrier'):
                dist.all_reduce(self._tok, group=self.group)

    # optional per-stage device timing (profile=True): CUDA events on the current stream, summed per stage name
    class _Span:
        def __init__(self, owner, name):
            self.o, self.name = owner, name

        def __enter__(self):
            if self.o.profile:
                self.a = torch.cuda.Event(enable_timing=True)
                self.a.record()

        def __exit__(self, *exc):
            if self.o.profile:
                b = torch.cuda.Event(enable_timing=True)
                b.record()
                self.o._spans.append((self.name, self.a, b))

    def _timed(self, name):
        return SlabPme._Span(self, name)

    def stage_times(self):
        """{stage: milliseconds} of the last evaluate() when profile=True (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self._spans:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    # ---- evaluation
    def evaluate(self, positions, box, pairs, Q_local, pol=None, tholes=None, mScales=None, pScales=None, U_init=None,
                 want_virial=True, maxiter=None, thresh=None):
        """Same contract as AtomBlockPme.evaluate: every rank passes the same full inputs and receives the same
        full outputs: dict(E, dpos, dbox, dQ_local, U, F, n_cycle, converged, scalars)."""
        from ._ctx import pairs_to_dev
        calc = self.calc
        self._setup()
        self._spans = []
        c, lib = calc._ctx, calc._ctx.lib
        p, sp = _lib.ptr, _lib.stream_ptr
        dev, dt = c.device, c.dtype
        n, nh = calc.n_atoms, (calc.lmax + 1) ** 2
        polz = pol is not None
        prep = calc._prep
        pos, box_d, Ql, mS = prep(positions).detach(), prep(box).detach(), prep(Q_local).detach(), prep(mScales).detach()
        pr = pairs_to_dev(pairs, dev)
        rows = partition_rows(int(pr.shape[0]), self.P)
        maxiter = settings.MAX_N_POL if maxiter is None else maxiter
        thresh = settings.POL_CONV if thresh is None else thresh
        fl = _lib.WANT_GRAD | (_lib.WANT_VIRIAL if want_virial else 0)
        vir = _lib.WANT_VIRIAL if want_virial else 0

        span_setup = self._timed('setup')
        span_setup.__enter__()
        for r in self.mine:
            _lib.check(lib.admp_set_box(self._ctxs[r].handle, sp(), p(box_d)))
        M = torch.empty((n, 10), dtype=dt, device=dev)
        _lib.check(lib.admp_frames_fwd(c.handle, sp(), p(pos), p(box_d), p(Ql), p(M), None, None))
        # spatial ownership (same rule as partition_atoms_by_slab, evaluated on the device in float64)
        sx = (pos.to(torch.float64) @ torch.linalg.inv(box_d.to(torch.float64)))[:, 0]
        sx = sx - torch.floor(sx)
        owner = torch.clamp((sx * self.P).to(torch.int64), max=self.P - 1)
        # per owned rank: spatial atom set (compact copies), pair-row slice, per-site atom block
        work = []
        for r in self.mine:
            idx = torch.nonzero(owner == r).reshape(-1)
            cnt = int(idx.numel())
            work.append(dict(r=r, ctx=self._ctxs[r], idx=idx, cnt=cnt, pos=pos.index_select(0, idx).contiguous(),
                             M=M.index_select(0, idx).contiguous(), pairs=pr[rows[r][0]:rows[r][0] + rows[r][1]].contiguous(),
                             nrows=rows[r][1], a0=self.atoms[r][0], ac=self.atoms[r][1],
                             Fr=torch.empty((cnt, 3), dtype=dt, device=dev)))
        span_setup.__exit__()
        scal = torch.zeros(_lib.S_COUNT, dtype=torch.float64, device=dev)
        U = F = None
        n_cycle, conv = 0, True
        if polz:
            polt, th, pS = prep(pol).detach(), prep(tholes).detach(), prep(pScales).detach()
            U = torch.zeros((n, 3), dtype=dt, device=dev) if U_init is None else prep(U_init).detach().clone()
            F = torch.zeros((n, 3), dtype=dt, device=dev)
            state = torch.zeros(8, dtype=torch.int32, device=dev)
            # The host drives the loop but never waits for the cycle it has just enqueued: cycle c + 1 is issued before the
            # decision of cycle c is read back (pinned snapshot + event). If cycle c ended the loop, the extra cycle is a
            # no-op for U and the SCF status (scf_decide_kernel's `ended` guard) and only recomputes the same potential.
            # Worth it while a cycle is short enough for the host's enqueue latency to show (measured: 32k waters on 2 GPUs
            # 86 -> 71 ms); on the largest meshes the one wasted cycle (1/31 of the GPU time) costs more than it hides.
            speculate = (calc.K1 * calc.K2 * calc.K3) / self.P < 2.5e8
            snaps = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(2)]
            events = [torch.cuda.Event() for _ in range(2)]
            pending = None
            final = None
            for c_it in range(maxiter + 2):
                scal.zero_()
                self._recip(work, U, scal, 0)
                F.zero_()
                with self._timed('gather'):
                    for wk in work:
                        Fr = wk['Fr'].zero_()
                        _lib.check(lib.admp_slab_gather(wk['ctx'].handle, sp(), p(wk['pos']), p(wk['M']), 10, 10, p(wk['U']), 1, 0, None, None,
                                                        10, p(Fr), p(scal), wk['cnt']))
                        F.index_add_(0, wk['idx'], Fr)
                with self._timed('pair'):
                    for wk in work:
                        if wk['nrows'] > 0:
                            # one row slice per context: its scale indices / cluster tiles are built in cycle 0 and reused
                            reuse = _lib.REUSE_PAIR_TILES if (len(work) == 1 and c_it > 0) else 0
                            _lib.check(lib.admp_pme_real(c.handle, sp(), p(pos), p(box_d), p(wk['pairs']), wk['nrows'], p(M), p(U), p(polt),
                                                         p(th), p(mS), p(pS), 1, reuse, None, None, p(F), None, None, p(scal)))
                with self._timed('allreduce_F'):
                    allreduce_sum_([F], self.group)      # also orders this cycle's gathers before the next slab_zero
                with self._timed('scf_step'):
                    _lib.check(lib.admp_scf_step(c.handle, sp(), p(M), p(U), p(polt), p(F), int(maxiter), float(thresh), vir, p(state),
                                                 p(scal)))
                    k = c_it % 2
                    snaps[k].copy_(state, non_blocking=True)
                    events[k].record()
                    if not speculate:
                        pending = k
                    if pending is not None:
                        events[pending].synchronize()
                        if not int(snaps[pending][5]):
                            final = snaps[pending].clone()
                            break
                    pending = k
            if final is None:
                events[pending].synchronize()
                final = snaps[pending].clone()
            n_cycle, conv = int(final[3]), bool(final[4])
            if want_virial:
                scal.zero_()
                self._recip(work, U, scal, vir)
        else:
            self._recip(work, None, scal, vir)
        # final evaluation at fixed U; phi of the last round trip sits in the slabs
        span_final = self._timed('final')
        span_final.__enter__()
        e_recip = scal[_lib.S_E_RECIP].clone()
        tk = scal[_lib.S_TK:_lib.S_TK + 6].clone()
        scal.zero_()
        scal[_lib.S_E_RECIP] = e_recip                     # this rank's share of the k sum: summed over ranks below
        scal[_lib.S_TK:_lib.S_TK + 6] = tk
        G = torch.zeros((n, 10), dtype=dt, device=dev)
        dpos = torch.zeros((n, 3), dtype=dt, device=dev)
        dQ = torch.zeros((n, nh), dtype=dt, device=dev)
        Fo = torch.zeros((n, 3), dtype=dt, device=dev) if polz else None
        for wk in work:
            cnt = wk['cnt']
            dr, Gr = torch.zeros((cnt, 3), dtype=dt, device=dev), torch.zeros((cnt, 10), dtype=dt, device=dev)
            Fr = torch.zeros((cnt, 3), dtype=dt, device=dev) if polz else None
            _lib.check(lib.admp_slab_gather(wk['ctx'].handle, sp(), p(wk['pos']), p(wk['M']), 10, 10, p(wk['U']) if polz else None, 0, fl,
                                            p(dr), p(Gr), 10, p(Fr), p(scal), cnt))
            dpos.index_add_(0, wk['idx'], dr)
            G.index_add_(0, wk['idx'], Gr)
            if polz:
                Fo.index_add_(0, wk['idx'], Fr)
            if wk['nrows'] > 0:
                reuse = _lib.REUSE_PAIR_TILES if (len(work) == 1 and polz) else 0       # built by the SCF cycles
                _lib.check(lib.admp_pme_real(c.handle, sp(), p(pos), p(box_d), p(wk['pairs']), wk['nrows'], p(M), p(U),
                                             p(polt) if polz else None, p(th) if polz else None, p(mS), p(pS) if polz else None, 0, fl | reuse,
                                             p(dpos), p(G), p(Fo), None, None, p(scal)))
            _lib.check(lib.admp_pme_self_range(c.handle, sp(), p(M), p(U), p(polt) if polz else None, fl, p(G), p(Fo), None, p(scal),
                                               wk['a0'], wk['ac']))
        allreduce_sum_([G], self.group)
        for wk in work:
            _lib.check(lib.admp_frames_bwd_range(c.handle, sp(), p(pos), p(Ql), p(G), p(dQ), p(dpos), p(scal), wk['a0'], wk['ac']))
        allreduce_sum_([dpos, dQ, Fo, scal], self.group)
        if want_virial:
            _lib.check(lib.admp_virial_finalize(c.handle, sp(), p(scal)))
        E = scal[_lib.S_E_REAL] + scal[_lib.S_E_RECIP] + scal[_lib.S_E_SELF] + scal[_lib.S_E_PEN]
        span_final.__exit__()
        return dict(E=E, dpos=dpos, dbox=scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3).clone(), dQ_local=dQ, U=U, F=Fo,
                    n_cycle=n_cycle, converged=conv, scalars=scal)

    def _recip(self, work, U, scal, vir):
        """zero | spread | Z,Y forward | fused X pass over peer memory | Y,Z inverse, with a cross-rank barrier
        between the stages; leaves phi = dE/dmesh in the slabs."""
        lib = self.calc._ctx.lib
        p, sp = _lib.ptr, _lib.stream_ptr
        with self._timed('zero'):
            for wk in work:
                wk['U'] = U.index_select(0, wk['idx']).contiguous() if U is not None else None
                _lib.check(lib.admp_slab_zero(wk['ctx'].handle, sp()))
        self._barrier()
        with self._timed('spread'):
            for wk in work:
                _lib.check(lib.admp_slab_spread(wk['ctx'].handle, sp(), p(wk['pos']), p(wk['M']), 10, 10, p(wk['U']), wk['cnt']))
        self._barrier()
        for phase in range(3):
            with self._timed(('fft_zy_fwd', 'fft_x_peer', 'fft_yz_inv')[phase]):
                for wk in work:
                    _lib.check(lib.admp_slab_fft(wk['ctx'].handle, sp(), phase, _lib.CK_COULOMB, vir, p(scal)))
            self._barrier()
