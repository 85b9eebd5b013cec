"""jax.ffi + jax.custom_vjp binding of libadmp_b200.so: `ADMPPmeForce` with the reference's surface (admp/pme.py:30-143) whose
energies come from the sm_100a kernels and whose derivatives are the analytic adjoints of the same call, so that
`jax.grad(get_energy, argnums=...)` / `jax.value_and_grad` work without tracing through the pair loop.

STATUS: UNTESTED - JAX cannot be installed in this repository's build image. It mirrors, call for call, the tested torch binding
(admp_b200/pme.py: `_PmeFunction` <-> `pme_energy` below). Needs libadmp_b200_xla.so built from admp_b200_xla.cc (same directory).
"""
import ctypes
import functools
import os

import jax
import jax.numpy as jnp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = ctypes.CDLL(os.path.join(HERE, '..', '..', 'admp_b200', 'lib', 'libadmp_b200.so'))
SHIM = ctypes.CDLL(os.path.join(HERE, 'libadmp_b200_xla.so'))
for _name in ('AdmpPmeEval', 'AdmpDispEval', 'AdmpTtPair', 'AdmpNblistBuild'):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(SHIM, _name)), platform='CUDA')

WANT_GRAD, WANT_VIRIAL, WANT_PGRAD, SCF = 1, 2, 4, 8
S_E, S_DBOX, S_DMSCALE, S_DPSCALE, S_COUNT = slice(0, 4), slice(4, 13), slice(28, 33), slice(33, 38), 48
POL_CONV, MAX_N_POL = 10.0, 30                                     # admp/settings.py:29-30


def _make_ctx(n_atoms, axis_type, axis_indices, cov_offsets, cov_index, cov_nbonds, kappa, K, lmax):
    """admp_ctx_create + admp_ctx_set_topology + admp_ctx_set_pme (see admp_b200/_ctx.py for the ctypes prototypes)."""
    h = ctypes.c_void_p()
    assert LIB.admp_ctx_create(ctypes.byref(h), 0, 0) == 0
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                 # noqa: E731
    assert LIB.admp_ctx_set_topology(h, n_atoms, p(axis_type), p(axis_indices), p(cov_offsets), p(cov_index), p(cov_nbonds)) == 0
    LIB.admp_ctx_set_pme.argtypes = [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_int] * 4
    assert LIB.admp_ctx_set_pme(h, float(kappa), int(K[0]), int(K[1]), int(K[2]), int(lmax)) == 0
    return h.value


def _pme_call(ctx, flags, positions, box, pairs, Q_local, U, pol, tholes, mScales, pScales):
    n, nh = positions.shape[0], Q_local.shape[1]
    f64 = lambda *s: jax.ShapeDtypeStruct(s, jnp.float64)          # noqa: E731
    out = jax.ffi.ffi_call('AdmpPmeEval', (f64(S_COUNT), f64(n, 3), f64(n, nh), f64(n, 3), f64(n, 3), f64(n), f64(n),
                                           jax.ShapeDtypeStruct((2,), jnp.int32)))(
        positions, box, pairs.astype(jnp.int32), Q_local, U, pol, tholes, mScales, pScales,
        ctx=np.int64(ctx), flags=np.int32(flags), maxiter=np.int32(MAX_N_POL), thresh=float(POL_CONV))
    return out      # scalars, dpos, dQ, F (= dE/dU), U_out, dpol, dtholes, scf


@functools.partial(jax.custom_vjp, nondiff_argnums=(0, 1, 4))
def pme_energy(ctx, do_scf, positions, box, pairs, Q_local, U, pol, tholes, mScales, pScales):
    flags = SCF if do_scf else 0
    return jnp.sum(_pme_call(ctx, flags, positions, box, pairs, Q_local, U, pol, tholes, mScales, pScales)[0][S_E])


def _fwd(ctx, do_scf, positions, box, pairs, Q_local, U, pol, tholes, mScales, pScales):
    flags = WANT_GRAD | WANT_VIRIAL | WANT_PGRAD | (SCF if do_scf else 0)
    s, dpos, dQ, F, _, dpol, dth, _ = _pme_call(ctx, flags, positions, box, pairs, Q_local, U, pol, tholes, mScales, pScales)
    # Hellmann-Feynman (admp/pme.py:83-85): with the SCF the gradient does not flow into U_init
    dU = jnp.zeros_like(F) if do_scf else F
    return jnp.sum(s[S_E]), (dpos, s[S_DBOX].reshape(3, 3), dQ, dU, dpol, dth, s[S_DMSCALE], s[S_DPSCALE])


def _bwd(ctx, do_scf, pairs, res, g):
    return tuple(g * r for r in res)


pme_energy.defvjp(_fwd, _bwd)


class ADMPPmeForce:
    """Drop-in for admp.pme.ADMPPmeForce (constructor, get_energy, get_forces, update_env; polarizable and not). The covalent map is
    taken in CSR form (cov_offsets, cov_index, cov_nbonds) - admp_b200.covalent.as_sparse converts the dense matrix."""

    def __init__(self, box, axis_type, axis_indices, covalent_csr, rc, ethresh, lmax, lpol=False):
        kappa = float(np.sqrt(-np.log(2 * ethresh)) / rc)           # admp/pme.py:146-172
        box = np.asarray(box)
        self.K = [int(np.ceil(2 * kappa * box[d, d] / 3 / ethresh ** 0.2)) for d in range(3)]
        self.kappa, self.lmax, self.lpol, self.n_atoms = kappa, lmax, lpol, len(axis_type)
        self._topo = (np.ascontiguousarray(axis_type, np.int32), np.ascontiguousarray(axis_indices, np.int32)) + tuple(covalent_csr)
        self.refresh_calculators()
        self.get_forces = jax.value_and_grad(self.get_energy)       # admp/pme.py:108

    def update_env(self, attr, val):
        if attr in ('K1', 'K2', 'K3'):
            self.K['K1 K2 K3'.split().index(attr)] = int(val)
        else:
            setattr(self, attr, val)
        self.refresh_calculators()

    def refresh_calculators(self):
        self._ctx = _make_ctx(self.n_atoms, *self._topo, self.kappa, self.K, self.lmax)

    def get_energy(self, positions, box, pairs, Q_local, *rest):
        if self.lpol:
            pol, tholes, mScales, pScales, dScales = rest[:5]
            U0 = rest[5] if len(rest) > 5 else jnp.zeros((self.n_atoms, 3))
            return pme_energy(self._ctx, True, positions, box, pairs, Q_local, U0, pol, tholes, mScales, pScales)
        e = jnp.zeros((0,))
        return pme_energy(self._ctx, False, positions, box, pairs, Q_local, jnp.zeros((0, 3)), e, e, rest[0], e)
