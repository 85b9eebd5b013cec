"""Module-level settings, same names and meaning as admp/settings.py:5-30.

PRECISION selects the arithmetic type of every kernel ('double' -> float64,
'single' -> float32, complex64 FFT).  DO_JIT / jit_condition are kept for source
compatibility: there is no tracing compiler here, kernels are ahead-of-time CUDA.
"""
PRECISION = 'double'
DO_JIT = True

# DEFAULT THRESHOLDS (admp/settings.py:29-30)
POL_CONV = 10.0   # gradient convergence thresh for induced dipoles
MAX_N_POL = 30    # maximum number of cycles for optimizing induced dipoles


# Induced-dipole solver (no counterpart in admp/settings.py). 'jacobi': the reference's iteration, cycle for cycle
# (admp/pme.py:132-138) - the default, and the only setting under which n_cycle / lconverg / U_ind reproduce the reference.
# 'pcg': conjugate gradients preconditioned with pol / DIELECTRIC on the same linear fixed point: same
# stopping rule (max|dE/dU| < POL_CONV over pol > 0.001, tested on the final U), MAX_N_POL bounds the CG iterations; it
# converges in fewer field evaluations and wherever the matrix is positive definite, also where the Jacobi iteration diverges.
SCF_SOLVER = 'jacobi'

# Convention of the k-space part of dE/dbox (no counterpart in admp/settings.py). 'natural': component i of a k-vector
# belongs to mesh axis i - the chain-rule-correct virial. 'reference': the reference's k table (admp/recip.py:339-341,
# meshgrid(kz, kx, ky)), which exchanges axes 0 and 1 in dk^2/dbox; on cubic cells with K1=K2=K3 this reproduces the
# diagonal of the reference's jax.grad(..., argnums=box) entry by entry. Read when a calculator (re)builds its plans
# (constructor / update_env / refresh_calculators). Energies, forces and all other gradients are unaffected.
KVEC_ORDER = 'natural'


def jit_condition(*args, **kwargs):
    """admp/settings.py:12-18: a decorator factory; a no-op here."""
    def deco(func):
        return func
    return deco


def torch_dtype(precision=None):
    import torch
    p = PRECISION if precision is None else precision
    if p == 'double':
        return torch.float64
    if p == 'single':
        return torch.float32
    raise ValueError("PRECISION must be 'double' or 'single', got %r" % (p,))
