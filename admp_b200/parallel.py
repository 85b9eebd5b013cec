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
def evaluate_frames(calc, frames, box, pairs, Q_local, pol=None, tholes=None, mScales=None, pScales=None,
                    rank=0, world=1, group=None, param_grads=True):
    """Frame-sharded evaluation with `calc` (an ADMPPmeForce). `frames`: sequence of (Na,3) arrays,
    `pairs`: one pair list or a callable frame_index -> pairs. Returns a dict with the local frame
    indices, their energies and dE/dpositions, and the frame-summed, rank-reduced parameter gradients."""
    mine = shard_frames(len(frames), rank, world)
    dev, dt = calc._ctx.device, calc._dtype
    prep = calc._prep
    box_d, Ql = prep(box), prep(Q_local)
    polz = calc.lpol
    rest = [prep(x) for x in ((pol, tholes, mScales, pScales) if polz else (mScales,))]
    flags = _lib.WANT_GRAD | (_lib.WANT_PGRAD if param_grads else 0)
    n, nh = calc.n_atoms, (calc.lmax + 1) ** 2
    acc = dict(dQ_local=torch.zeros((n, nh), dtype=torch.float64, device=dev),
               dmScales=torch.zeros(5, dtype=torch.float64, device=dev))
    if polz:
        acc.update(dpScales=torch.zeros(5, dtype=torch.float64, device=dev),
                   dtholes=torch.zeros(n, dtype=torch.float64, device=dev),
                   dpol=torch.zeros(n, dtype=torch.float64, device=dev))
    energies, grads = [], []
    zeroU = torch.zeros((n, 3), dtype=dt, device=dev) if polz else None
    for f in mine:
        pr = pairs(f) if callable(pairs) else pairs
        from ._ctx import pairs_to_dev
        pr = pairs_to_dev(pr, dev)
        if polz:
            r = calc._eval(prep(frames[f]), box_d, pr, Ql, zeroU, rest[0], rest[1], rest[2], rest[3], flags, True, cache_scf=False)
        else:
            r = calc._eval(prep(frames[f]), box_d, pr, Ql, None, None, None, rest[0], None, flags, False)
        energies.append(r.energy)
        grads.append(r.dpos)
        if param_grads:
            acc['dQ_local'] += r.dQ
            acc['dmScales'] += r.scalars[_lib.S_DMSCALE:_lib.S_DMSCALE + 5]
            if polz:
                acc['dpScales'] += r.scalars[_lib.S_DPSCALE:_lib.S_DPSCALE + 5]
                acc['dtholes'] += r.dtholes
                acc['dpol'] += r.dpol
    if param_grads and world > 1:
        allreduce_sum_(list(acc.values()), group)
    return dict(frames=mine, energies=torch.stack(energies) if energies else torch.zeros(0, dtype=torch.float64, device=dev),
                dpos=grads, param_grads=acc if param_grads else None)


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
