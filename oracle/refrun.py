"""Run the UNMODIFIED reference sources under the jax API shim (TEST INFRASTRUCTURE).

``load_reference()`` imports ``admp.{settings,multipole,spatial,pairwise,pme,recip,disp_pme}`` from
where they lie under /root/reference, with ``oracle/jaxshim`` standing in for jax (README there).
It needs /root/reference, so it is only usable in the build container: the GPU box sees its outputs
as committed fixtures (tests/golden/ref_*.npz, written by tests/golden/make_reference_goldens.py).

Nothing of the reference is copied; the modules are executed from their own files. After loading,
``jax`` / ``admp`` are removed from ``sys.modules`` and ``sys.path`` again so that the rest of the
process (pytest, torch, anything probing for a real jax) is unaffected.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = '/root/reference'
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'jaxshim')
_MODULES = ('settings', 'multipole', 'spatial', 'pairwise', 'pme', 'recip', 'disp_pme')
_cache = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'admp', 'pme.py'))


def load_reference():
    """-> namespace with .settings .multipole .spatial .pairwise .pme .recip .disp_pme (reference
    modules) and .jnp / .jax (the shim they run on)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError('%s/admp not present (the reference only exists in the build container)' % REFERENCE_ROOT)
    saved = {k: v for k, v in sys.modules.items() if k == 'jax' or k.startswith('jax.') or k == 'admp' or k.startswith('admp.')}
    for k in saved:
        del sys.modules[k]
    sys.path[:0] = [SHIM_ROOT, REFERENCE_ROOT]
    try:
        ns = types.SimpleNamespace()
        ns.jax = importlib.import_module('jax')
        assert getattr(ns.jax, '__shim__', False), 'a real jax is importable: use it instead of the shim'
        ns.jnp = importlib.import_module('jax.numpy')
        for m in _MODULES:
            mod = importlib.import_module('admp.' + m)
            assert mod.__file__.startswith(REFERENCE_ROOT), mod.__file__
            setattr(ns, m, mod)
    finally:
        sys.path.remove(SHIM_ROOT)
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == 'jax' or k.startswith('jax.') or k == 'admp' or k.startswith('admp.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns


def A(x):
    """torch tensor / ndarray -> shim array (float64 / int64)."""
    ns = load_reference()
    arr_t = ns.jax.Array
    if isinstance(x, arr_t):
        return x
    if isinstance(x, torch.Tensor):
        return arr_t(x.detach().clone())
    return arr_t(torch.as_tensor(np.asarray(x)).clone())


def T(x):
    """shim array (or tuple of) -> detached torch tensor(s)."""
    if isinstance(x, (tuple, list)):
        return type(x)(T(u) for u in x)
    return x.t.detach() if hasattr(x, 't') else x
