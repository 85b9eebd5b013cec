"""Thin Python owner of an ``admp_ctx`` plus tensor marshalling helpers."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import settings
from .covalent import as_sparse


def device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(x, dtype, dev=None):
    """Array-like / tensor -> contiguous device tensor of ``dtype`` (keeps the autograd link)."""
    dev = device() if dev is None else dev
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.device != dev or t.dtype != dtype:
        t = t.to(device=dev, dtype=dtype)
    return t.contiguous()


def pairs_to_dev(pairs, dev=None):
    dev = device() if dev is None else dev
    if isinstance(pairs, torch.Tensor):
        t = pairs
    else:
        t = torch.as_tensor(np.asarray(pairs))
    if t.dim() != 2 or t.shape[1] != 2:
        raise ValueError('pairs must have shape (Np, 2)')
    return t.to(device=dev, dtype=torch.int32).contiguous()


class Context:
    """One ``admp_ctx`` (cuFFT plans, workspaces, topology) per calculator; not re-entrant,
    like the reference calculators (admp/pme.py:82 mutates self)."""

    def __init__(self, precision=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.dtype = settings.torch_dtype(precision)
        self.device = device()
        h = ctypes.c_void_p()
        _lib.check(self.lib.admp_ctx_create(ctypes.byref(h), self.device.index,
                                            _lib.F64 if self.dtype == torch.float64 else _lib.F32))
        self.handle = h
        self.n_atoms = 0

    def set_pme(self, kappa, K1, K2, K3, lmax):
        _lib.check(self.lib.admp_ctx_set_pme(self.handle, float(kappa), int(K1), int(K2), int(K3), int(lmax)))
        if settings.KVEC_ORDER not in ('natural', 'reference'):
            raise ValueError("settings.KVEC_ORDER must be 'natural' or 'reference'")
        _lib.check(self.lib.admp_ctx_set_kvec_order(self.handle, 1 if settings.KVEC_ORDER == 'reference' else 0))

    def set_topology(self, n_atoms, axis_type=None, axis_indices=None, covalent_map=None):
        at = ai = off = idx = nb = None
        keep = []
        if axis_type is not None and axis_indices is not None:
            at = np.ascontiguousarray(np.asarray(axis_type), dtype=np.int32)
            ai = np.ascontiguousarray(np.asarray(axis_indices), dtype=np.int32).reshape(-1, 3)
            if at.shape[0] != n_atoms or ai.shape[0] != n_atoms:
                raise ValueError('axis_type / axis_indices must have n_atoms rows')
        if covalent_map is not None:
            sp = as_sparse(covalent_map)
            if sp.n_atoms != n_atoms:
                raise ValueError('covalent_map is %d x %d but the system has %d atoms' % (sp.n_atoms, sp.n_atoms, n_atoms))
            off, idx, nb = sp.offsets, sp.index, sp.nbonds
        keep = [at, ai, off, idx, nb]
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self.lib.admp_ctx_set_topology(self.handle, int(n_atoms), p(at), p(ai), p(off), p(idx), p(nb)))
        self.n_atoms = int(n_atoms)
        del keep

    def mesh_view(self, K):
        """Zero-copy torch view (K1, K2, K3) of the context's real mesh (the buffer admp_pme_spread_range
        accumulates into and the FFT passes read): lets torch.distributed all-reduce it in place."""
        self.lib.admp_ctx_buffer.restype = ctypes.c_void_p
        ptr = self.lib.admp_ctx_buffer(self.handle, 0)
        if not ptr:
            raise _lib.AdmpLibraryError('admp_ctx_set_pme has not been called')

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = dict(shape=tuple(int(k) for k in K), typestr='<f8' if self.dtype == torch.float64 else '<f4',
                                            data=(int(ptr), False), version=2)
        t = torch.as_tensor(raw, device=self.device)
        t._admp_ctx_keepalive = self            # the view must not outlive the context's allocation
        return t

    @property
    def workspace_bytes(self):
        return int(self.lib.admp_ctx_workspace_bytes(self.handle))

    @property
    def scf_graph_active(self):
        return bool(self.lib.admp_ctx_scf_graph_active(self.handle))

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle.value:
            self.lib.admp_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
