"""Sparse covalent map.

The reference passes a dense Na x Na integer matrix ``covalent_map`` (bond count, 0 =
not bonded; admp/parser.py:462-476, admp/api.py:24-42) and indexes it per pair
(admp/pme.py:681).  That is 4.9 TB at 786k atoms, so the kernels take the same
information as a CSR list of bonded partners per atom.  Dense matrices are still
accepted at the API (``as_sparse``) for drop-in use on small systems.
"""
import numpy as np


class SparseCovalentMap:
    """CSR: for atom i, partners ``index[offsets[i]:offsets[i+1]]`` with bond counts ``nbonds``."""

    def __init__(self, n_atoms, offsets, index, nbonds):
        self.n_atoms = int(n_atoms)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        self.index = np.ascontiguousarray(index, dtype=np.int32)
        self.nbonds = np.ascontiguousarray(nbonds, dtype=np.int8)
        if self.offsets.shape != (self.n_atoms + 1,):
            raise ValueError('offsets must have n_atoms + 1 entries')
        if self.index.shape != self.nbonds.shape or int(self.offsets[-1]) != self.index.shape[0]:
            raise ValueError('inconsistent CSR arrays')

    @property
    def shape(self):
        """Mimics the dense matrix (``covalent_map.shape[0]`` is read at admp/pme.py:52)."""
        return (self.n_atoms, self.n_atoms)

    @classmethod
    def from_pairs(cls, n_atoms, i, j, nbonds):
        i = np.asarray(i, dtype=np.int64)
        j = np.asarray(j, dtype=np.int64)
        nb = np.asarray(nbonds)
        order = np.lexsort((j, i))
        i, j, nb = i[order], j[order], nb[order]
        offsets = np.searchsorted(i, np.arange(n_atoms + 1))
        return cls(n_atoms, offsets, j, nb)

    @classmethod
    def from_dense(cls, dense):
        dense = np.asarray(dense)
        if dense.ndim != 2 or dense.shape[0] != dense.shape[1]:
            raise ValueError('covalent_map must be a square matrix')
        i, j = np.nonzero(dense)
        return cls.from_pairs(dense.shape[0], i, j, dense[i, j])

    def dense(self):
        m = np.zeros((self.n_atoms, self.n_atoms), dtype=np.int64)
        rows = np.repeat(np.arange(self.n_atoms), np.diff(self.offsets))
        m[rows, self.index] = self.nbonds
        return m


def as_sparse(covalent_map):
    """Accept a SparseCovalentMap, anything with CSR-like (ci, cj, cn) arrays, or a dense matrix."""
    if isinstance(covalent_map, SparseCovalentMap):
        return covalent_map
    if hasattr(covalent_map, 'ci') and hasattr(covalent_map, 'cj') and hasattr(covalent_map, 'cn'):
        return SparseCovalentMap.from_pairs(covalent_map.shape[0], covalent_map.ci, covalent_map.cj, covalent_map.cn)
    if hasattr(covalent_map, 'detach'):
        covalent_map = covalent_map.detach().cpu().numpy()
    return SparseCovalentMap.from_dense(np.asarray(covalent_map))
