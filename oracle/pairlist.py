"""Oracle (test infrastructure): neighbour pair set.

The reference has no neighbour list of its own: callers use
``jax_md.partition.neighbor_list(space.periodic_general(box)[0], box, rc, 0,
format=OrderedSparse)`` (examples/water_1024/run_admp.py:109-112; jax-md is an
UNPINNED third-party dependency absent from /root/reference => "parity unpinned":
no reference fixture holds a pair set).  What is restated here is the published
predicate: i<j, minimum-image displacement through fractional coordinates,
``d.d < rc^2`` strict.

Canonical arithmetic (shared bit-for-bit with the CUDA builder, all float64, no
fused multiply-add):
    s   = r * (1/L)            per atom and per dimension (orthorhombic boxes)
    t   = (s_i - s_j) + 0.5
    ds  = (t - floor(t)) - 0.5
    d   = ds * L
    d2  = (d_x*d_x + d_y*d_y) + d_z*d_z
    keep iff d2 < rc*rc
Output rows are (i, j) with i < j sorted lexicographically, padded with (N, N).
"""
import numpy as np


def _d2_block(si, sj, L):
    t = (si[:, None, :] - sj[None, :, :]) + 0.5
    ds = (t - np.floor(t)) - 0.5
    d = ds * L
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def build_pairs(positions, box, rc, capacity=None, block=2048):
    """O(N^2) blocked reference builder.  Returns (pairs (cap,2) int32, n_pairs)."""
    r = np.asarray(positions, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    L = np.array([box[0, 0], box[1, 1], box[2, 2]])
    if np.abs(box - np.diag(L)).max() != 0.0:
        raise NotImplementedError('pair builder: orthorhombic boxes only')
    n = r.shape[0]
    s = r * (1.0 / L)
    rc2 = float(rc) * float(rc)
    out = []
    for a in range(0, n, block):
        for b in range(a, n, block):
            d2 = _d2_block(s[a:a + block], s[b:b + block], L)
            ii, jj = np.nonzero(d2 < rc2)
            ii = ii + a
            jj = jj + b
            keep = ii < jj
            out.append(np.stack([ii[keep], jj[keep]], axis=1))
    pairs = np.concatenate(out, axis=0) if out else np.zeros((0, 2), dtype=np.int64)
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    pairs = pairs[order].astype(np.int32)
    npairs = pairs.shape[0]
    if capacity is None:
        capacity = npairs
    if npairs > capacity:
        raise OverflowError('pair capacity exceeded')
    pad = np.full((capacity - npairs, 2), n, dtype=np.int32)
    return np.concatenate([pairs, pad], axis=0), npairs
