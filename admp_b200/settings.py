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
