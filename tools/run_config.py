#!/usr/bin/env python
"""Runs one BASELINE config on one GPU and prints energy, SCF status, timing and replica invariance
against the base cell:   python tools/run_config.py C1|C2|C3|C5 [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                  # noqa: E402
import torch                                        # noqa: E402
from admp_b200 import _lib, workloads               # noqa: E402
from admp_b200.pme import ADMPPmeForce              # noqa: E402
from admp_b200.disp_pme import ADMPDispPmeForce     # noqa: E402
from admp_b200.neighbor import neighbor_list        # noqa: E402

CFG = {'C1': ((1, 1, 1), False), 'C2': ((1, 1, 1), True), 'C3': ((2, 4, 4), True), 'C5': ((4, 8, 8), True),
       'C2x2': ((2, 1, 1), True)}


def run(name, steps=3, disp=False):
    reps, pol = CFG[name]
    w = workloads.water_box(reps, polarizable=pol)
    t0 = time.perf_counter()
    calc = ADMPPmeForce(workloads.water_box((1, 1, 1)).box, w.axis_type, w.axis_indices, w.covalent_map, w.rc, w.ethresh, 2, lpol=pol)
    calc.update_env('kappa', w.kappa)
    for d in range(3):
        calc.update_env('K%d' % (d + 1), w.K[d])
    nbr = neighbor_list(w.box, w.rc).allocate(w.positions)
    torch.cuda.synchronize()
    print('%s: %d atoms, mesh %s, %d pairs, set-up %.2f s, workspace %.2f GB, fft backend %d' % (
        name, w.n_atoms, w.K, nbr.n_pairs, time.perf_counter() - t0, calc._ctx.workspace_bytes / 1e9,
        calc._ctx.lib.admp_ctx_fft_backend(calc._ctx.handle)))
    args = [calc._prep(x) for x in (w.positions, w.box, w.Q_local)]
    rest = [calc._prep(x) for x in ((w.pol, w.tholes, w.mScales, w.pScales) if pol else (w.mScales,))]
    fl = _lib.WANT_GRAD | _lib.WANT_VIRIAL
    ts = []
    for k in range(steps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if pol:
            r = calc._eval(args[0], args[1], nbr.pairs, args[2], None, rest[0], rest[1], rest[2], rest[3], fl, True)
        else:
            r = calc._eval(args[0], args[1], nbr.pairs, args[2], None, None, None, rest[0], None, fl, False)
        b.record()
        b.synchronize()
        if k:
            ts.append(a.elapsed_time(b))
    E = r.energy.item()
    scf = r.scf.cpu().tolist() if pol else None
    print('%s: E = %.6f kJ/mol, scf [n_cycle, converged] = %s, %.3f ms/eval (%.2f evals/s), max|dE/dr| %.4f' % (
        name, E, scf, np.mean(ts), 1e3 / np.mean(ts), r.dpos.abs().max().item()))
    out = dict(E=E, F=r.dpos[:3072].cpu(), U=(r.U[:3072].cpu() if pol else None), nrep=int(np.prod(reps)))
    if disp:
        dcalc = ADMPDispPmeForce(workloads.water_box((1, 1, 1)).box, w.covalent_map, w.rc, w.ethresh, 10)
        dcalc.update_env('kappa', w.kappa)
        for d in range(3):
            dcalc.update_env('K%d' % (d + 1), w.K[d])
        Ed, Fd = dcalc.get_forces(w.positions, w.box, nbr.pairs, w.c_list, w.mScales)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            Ed, Fd = dcalc.get_forces(w.positions, w.box, nbr.pairs, w.c_list, w.mScales)
        torch.cuda.synchronize()
        print('%s: dispersion PME E = %.6f, %.3f ms/eval' % (name, Ed.item(), 1e3 * (time.perf_counter() - t0) / steps))
        out['Ed'] = Ed.item()
    return out


if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    res = run(name, steps, disp=(name == 'C3'))
    if name not in ('C1', 'C2'):
        base = run('C2', 2, disp=(name == 'C3'))
        n = res['nrep']
        print('replica invariance: E/(n E_C2) - 1 = %.3e ; max|F - F_C2|/max|F| = %.3e ; max|U - U_C2|/max|U| = %.3e' % (
            res['E'] / (n * base['E']) - 1, (res['F'] - base['F']).abs().max().item() / base['F'].abs().max().item(),
            (res['U'] - base['U']).abs().max().item() / base['U'].abs().max().item()))
        if 'Ed' in res:
            print('dispersion replica invariance: %.3e' % (res['Ed'] / (n * base['Ed']) - 1))
