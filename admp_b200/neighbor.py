"""GPU neighbour list - the replacement for the ``jax_md.partition.neighbor_list(...,
format=OrderedSparse)`` call every reference script makes
(examples/water_1024/run_admp.py:109-112):

    neighbor_list_fn = neighbor_list(box, rc, 0)
    nbr = neighbor_list_fn.allocate(positions)
    pairs = nbr.idx.T            # (capacity, 2) int32, i<j rows then (N, N) padding

The pair SET is bit-exact w.r.t. oracle/pairlist.py (same float64 operation order).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._ctx import Context, to_dev


class NeighborList:
    """Result object shaped like jax_md's: ``idx`` (2, capacity), ``did_buffer_overflow``, ``update``."""

    def __init__(self, fn, pairs, info, reference_positions):
        self._fn = fn
        self._pairs = pairs
        self._info = info
        self.reference_position = reference_positions

    @property
    def idx(self):
        return self._pairs.T

    @property
    def pairs(self):
        return self._pairs

    @property
    def n_pairs(self):
        return int(self._info[0].item())

    @property
    def did_buffer_overflow(self):
        return bool(self._info[1].item())

    def update(self, positions):
        return self._fn.update(positions, self)


class NeighborListFn:
    def __init__(self, box, r_cutoff, dr_threshold=0.0, capacity_multiplier=1.25):
        self.box = box
        self.rc = float(r_cutoff)
        self.dr_threshold = float(dr_threshold)
        self.capacity_multiplier = float(capacity_multiplier)
        self._ctx = Context()
        # host and device copies of the box, made once: a build then needs no device-to-host read and no host sync
        hb = box.detach().cpu().numpy() if isinstance(box, torch.Tensor) else np.asarray(box)
        self._box_host = np.ascontiguousarray(hb, dtype=np.float64).reshape(3, 3)
        self._box_dev = to_dev(self._box_host, self._ctx.dtype, self._ctx.device)

    def _build(self, positions, capacity):
        cx = self._ctx
        pos = to_dev(positions, cx.dtype, cx.device).detach()
        n = int(pos.shape[0])
        pairs = torch.empty((capacity, 2), dtype=torch.int32, device=cx.device)
        info = torch.zeros(2, dtype=torch.int32, device=cx.device)
        _lib.check(cx.lib.admp_nblist_build_hostbox(cx.handle, _lib.stream_ptr(), _lib.ptr(pos), _lib.ptr(self._box_dev),
                                                    self._box_host.ctypes.data_as(ctypes.c_void_p), n, self.rc + self.dr_threshold,
                                                    _lib.ptr(pairs), int(capacity), _lib.ptr(info)))
        return NeighborList(self, pairs, info, pos)

    def allocate(self, positions, extra_capacity=0):
        """Two passes: count with a minimal buffer, then allocate capacity_multiplier * count."""
        probe = self._build(positions, 1)
        count = probe.n_pairs
        capacity = int(np.ceil(count * self.capacity_multiplier)) + int(extra_capacity)
        return self._build(positions, max(capacity, 1))

    def update(self, positions, nbr):
        """jax_md semantics: with a skin (dr_threshold > 0) the list is kept while no atom has moved further than
        dr_threshold / 2 from the positions it was built on (it was built with rc + dr_threshold, so it still holds
        every pair inside rc; like the reference, the energy functions evaluate every listed pair); otherwise it is
        rebuilt into a buffer of the same capacity (check ``did_buffer_overflow``). The test costs one device
        reduction and a scalar read."""
        if self.dr_threshold > 0.0:
            cx = self._ctx
            pos = to_dev(positions, cx.dtype, cx.device).detach()
            ref = nbr.reference_position
            if ref.shape == pos.shape:
                moved = (pos - ref).pow(2).sum(dim=1).max()
                if float(moved.item()) < (0.5 * self.dr_threshold) ** 2:
                    return nbr
        return self._build(positions, int(nbr.pairs.shape[0]))


def neighbor_list(box, r_cutoff, dr_threshold=0.0, capacity_multiplier=1.25, *extra, **_ignored):
    """Factory with jax_md's argument order minus the displacement function (the periodic general displacement is
    implied by ``box``). The reference scripts' own call, ``partition.neighbor_list(displacement_fn, box, rc, 0,
    format=partition.OrderedSparse)`` (examples/water_1024/run_admp.py:110-111), is accepted as well: a callable first
    argument is dropped."""
    if callable(box):
        box, r_cutoff, dr_threshold = r_cutoff, dr_threshold, (capacity_multiplier if extra or capacity_multiplier != 1.25 else 0.0)
        capacity_multiplier = extra[0] if extra else 1.25
    return NeighborListFn(box, r_cutoff, dr_threshold, capacity_multiplier)


class _Space:
    """``jax_md.space`` stand-in for the two lines the reference scripts use (run_admp.py:109)."""

    @staticmethod
    def periodic_general(box, fractional_coordinates=False):
        import torch as _torch
        b = box if isinstance(box, _torch.Tensor) else _torch.as_tensor(np.asarray(box, dtype=np.float64))

        def displacement_fn(ra, rb):
            from .spatial import pbc_shift
            return pbc_shift((ra - rb).reshape(-1, 3), b, _torch.linalg.inv(b)).reshape(ra.shape)

        def shift_fn(r, dr):
            return r + dr
        return displacement_fn, shift_fn


class _Partition:
    """``jax_md.partition`` stand-in: ``neighbor_list`` and the format marker."""
    OrderedSparse = 'OrderedSparse'
    neighbor_list = staticmethod(neighbor_list)


space = _Space()
partition = _Partition()
