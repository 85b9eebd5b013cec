"""jax.scipy subset of the shim (test infrastructure)."""
from . import special   # noqa: F401
