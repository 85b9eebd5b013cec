"""Dispersion PME calculator - drop-in surface of admp/disp_pme.py:20-77.

E = sum_pairs sum_{p=6,8,10} (m + g_p(kappa^2 r^2) - 1) c_i^p c_j^p / r^p
    + three lmax=0 reciprocal passes (Ck_6/8/10, gamma point kept) + self term.
The physical dispersion energy is -E (admp/api.py:199).  Differentiable inputs:
positions, box, c_list, mScales.
"""
import torch

from . import _lib
from ._ctx import Context, to_dev, pairs_to_dev
from .pme import setup_ewald_parameters


class _DispFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, calc, pairs, positions, box, c_list, mScales):
        n = ctx.needs_input_grad
        flags = 0
        if n[2] or n[3] or n[4] or n[5]:
            flags |= _lib.WANT_GRAD
        if n[3]:
            flags |= _lib.WANT_VIRIAL
        if n[4] or n[5]:
            flags |= _lib.WANT_PGRAD
        scal, dpos, dc = calc._eval(positions, box, pairs, c_list, mScales, flags)
        ctx.saved = (scal, dpos, dc)
        ctx.dtype = positions.dtype
        return (scal[_lib.S_E_REAL] + scal[_lib.S_E_RECIP] + scal[_lib.S_E_SELF]).to(positions.dtype)

    @staticmethod
    def backward(ctx, g):
        scal, dpos, dc = ctx.saved
        n, dt = ctx.needs_input_grad, ctx.dtype
        return (None, None,
                g * dpos if n[2] else None,
                (g * scal[_lib.S_DBOX:_lib.S_DBOX + 9].reshape(3, 3)).to(dt) if n[3] else None,
                g * dc if n[4] else None,
                (g * scal[_lib.S_DMSCALE:_lib.S_DMSCALE + 5]).to(dt) if n[5] else None)


class ADMPDispPmeForce:
    '''
    This is a convenient wrapper for dispersion PME calculations
    (same constructor and attributes as admp/disp_pme.py:27-41)
    '''

    def __init__(self, box, covalent_map, rc, ethresh, pmax):
        self.covalent_map = covalent_map
        self.rc = rc
        self.ethresh = ethresh
        self.pmax = pmax
        kappa, K1, K2, K3 = setup_ewald_parameters(rc, ethresh, box)
        self.kappa = kappa
        self.K1 = K1
        self.K2 = K2
        self.K3 = K3
        self.pme_order = 6
        self.n_atoms = int(covalent_map.shape[0])
        self._ctx = Context()
        self._dtype = self._ctx.dtype
        self._topology_set = False
        self.refresh_calculators()

    def update_env(self, attr, val):
        '''admp/disp_pme.py:53-58'''
        setattr(self, attr, val)
        if attr == 'covalent_map':
            self._topology_set = False
        self.refresh_calculators()

    def refresh_calculators(self):
        '''admp/disp_pme.py:61-77'''
        if self.pmax not in (6, 8, 10):
            raise ValueError('pmax must be 6, 8 or 10')
        self._ctx.set_pme(self.kappa, self.K1, self.K2, self.K3, 0)
        if not self._topology_set:
            self._ctx.set_topology(self.n_atoms, None, None, self.covalent_map)
            self._topology_set = True

    def _eval(self, positions, box, pairs, c_list, mScales, flags):
        c = self._ctx
        n, dt, dev = self.n_atoms, self._dtype, c.device
        if positions.shape != (n, 3) or c_list.shape != (n, 3):
            raise ValueError('positions and c_list must be (%d, 3); c_list columns are C6, C8, C10' % n)
        if tuple(box.shape) != (3, 3) or mScales.shape != (5,):
            raise ValueError('box must be (3, 3) and mScales must hold exactly 5 entries')
        scal = torch.empty(_lib.S_COUNT, dtype=torch.float64, device=dev)
        dpos = torch.empty((n, 3), dtype=dt, device=dev) if flags & _lib.WANT_GRAD else None
        dc = torch.empty((n, 3), dtype=dt, device=dev) if flags & _lib.WANT_PGRAD else None
        p = _lib.ptr
        _lib.check(c.lib.admp_disp_eval(c.handle, _lib.stream_ptr(), p(positions), p(box), p(pairs), int(pairs.shape[0]),
                                        p(c_list), p(mScales), int(self.pmax), flags, p(scal), p(dpos), p(dc)))
        return scal, dpos, dc

    def _prep(self, x):
        return to_dev(x, self._dtype, self._ctx.device)

    def _prep_c(self, c_list):
        """The reference's contract is Na x (pmax-4)/2 columns (admp/disp_pme.py:92); the kernels stride by 3. Narrower
        inputs are zero-padded (differentiably: autograd slices the gradient back to the caller's width)."""
        c = self._prep(c_list)
        need = (int(self.pmax) - 4) // 2
        if c.dim() != 2 or c.shape[0] != self.n_atoms or not (need <= c.shape[1] <= 3):
            raise ValueError('c_list must be (%d, k) with %d <= k <= 3 for pmax = %d' % (self.n_atoms, need, self.pmax))
        if c.shape[1] < 3:
            c = torch.nn.functional.pad(c, (0, 3 - c.shape[1]))
        return c.contiguous()

    def get_energy(self, positions, box, pairs, c_list, mScales):
        """admp/disp_pme.py:44-50"""
        positions, box, mScales = map(self._prep, (positions, box, mScales))
        c_list = self._prep_c(c_list)
        return _DispFunction.apply(self, pairs_to_dev(pairs, self._ctx.device), positions, box, c_list, mScales)

    def get_forces(self, positions, box, pairs, c_list, mScales):
        """value_and_grad(get_energy): (E, +dE/dpositions)  (admp/disp_pme.py:76)"""
        positions, box, mScales = (self._prep(x).detach() for x in (positions, box, mScales))
        c_list = self._prep_c(c_list).detach()
        scal, dpos, _ = self._eval(positions, box, pairs_to_dev(pairs, self._ctx.device), c_list, mScales, _lib.WANT_GRAD)
        return (scal[_lib.S_E_REAL] + scal[_lib.S_E_RECIP] + scal[_lib.S_E_SELF]).to(self._dtype), dpos

    def generate_get_energy(self):
        """admp/disp_pme.py:44-50"""
        self.refresh_calculators()
        return self.get_energy


# ---------------------------------------------------------------------- module-level functions of admp/disp_pme.py
def g_p(x2, pmax):
    """admp/disp_pme.py:219-252: g_p(x^2) = exp(-x^2) sum_{k < p/2} x^(2k) / k!, stacked for p = 6, 8, 10 (<= pmax)."""
    x2 = x2 if isinstance(x2, torch.Tensor) else torch.as_tensor(x2, dtype=torch.float64)
    x4 = x2 * x2
    g = [1 + x2 + 0.5 * x4]
    if pmax >= 8:
        g.append(g[0] + x4 * x2 / 6)
    if pmax >= 10:
        g.append(g[1] + x4 * x4 / 24)
    return torch.stack(g) * torch.exp(-x2)


def disp_pme_self(c_list, kappa, pmax):
    """admp/disp_pme.py:255-279: -kappa^6/12 sum c6^2 - kappa^8/48 sum c8^2 - kappa^10/240 sum c10^2."""
    c = c_list if isinstance(c_list, torch.Tensor) else torch.as_tensor(c_list, dtype=torch.float64)
    E = -kappa ** 6 / 12 * torch.sum(c[:, 0] ** 2)
    if pmax >= 8:
        E = E - kappa ** 8 / 48 * torch.sum(c[:, 1] ** 2)
    if pmax >= 10:
        E = E - kappa ** 10 / 240 * torch.sum(c[:, 2] ** 2)
    return E


_disp_cache = {}


def _disp_calc(box, covalent_map, kappa, K1, K2, K3, pmax):
    from . import settings
    key = (id(covalent_map), int(K1), int(K2), int(K3), int(pmax), settings.PRECISION)
    hit = _disp_cache.get(key)
    calc = hit[0] if (hit is not None and hit[1] is covalent_map) else None    # id() key: identity-checked
    if calc is None:
        import numpy as np
        calc = ADMPDispPmeForce(np.eye(3) * 20.0, covalent_map, 4.0, 1e-4, pmax)
        calc.K1, calc.K2, calc.K3 = int(K1), int(K2), int(K3)
        calc.kappa = float(kappa)
        calc.refresh_calculators()
        _disp_cache[key] = (calc, covalent_map)
    if calc.kappa != float(kappa):
        calc.update_env('kappa', float(kappa))
    return calc


def energy_disp_pme(positions, box, pairs, c_list, mScales, covalent_map, kappa, K1, K2, K3, pmax,
                    recip_fn6=None, recip_fn8=None, recip_fn10=None):
    """admp/disp_pme.py:80-123, the top-level dispersion-PME energy (the recip_fn* arguments are accepted for
    signature compatibility; the fused kernels of the calculator are used). Differentiable like get_energy."""
    return _disp_calc(box, covalent_map, kappa, K1, K2, K3, pmax).get_energy(positions, box, pairs, c_list, mScales)


def disp_pme_real(positions, box, pairs, c_list, mScales, covalent_map, kappa, pmax):
    """admp/disp_pme.py:126-178: the real-space part alone. The pair kernel is fused into admp_disp_eval, so this runs
    one evaluation on the smallest mesh and returns its real-space slot (no gradient; use ADMPDispPmeForce for those)."""
    calc = _disp_calc(box, covalent_map, kappa, 6, 6, 6, pmax)
    args = [calc._prep(x).detach() for x in (positions, box, c_list, mScales)]
    scal, _, _ = calc._eval(args[0], args[1], pairs_to_dev(pairs, calc._ctx.device), args[2], args[3], 0)
    return scal[_lib.S_E_REAL].to(calc._dtype)
