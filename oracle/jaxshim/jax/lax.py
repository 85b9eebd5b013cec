"""jax.lax subset of the shim (test infrastructure)."""
from ._array import Array


def stop_gradient(x):
    if isinstance(x, Array):
        return Array(x.t.detach())
    return x
