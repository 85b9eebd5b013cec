"""Regenerates tests/golden/ref_<case>.npz: outputs of the UNMODIFIED reference sources
(/root/reference/admp/*.py executed under oracle/jaxshim, see oracle/refrun.py) on the fixed inputs of
oracle/refcases.py.  Only runnable where /root/reference exists (the build container).

    python tests/golden/make_reference_goldens.py            # small cases (seconds)
    python tests/golden/make_reference_goldens.py c1 c2      # full-size BASELINE configs (minutes)

Every array is what the reference's own functions return:
  energy_pme (pme.py:176) and its terms pme_real (:628) / pme_recip (recip.py:394) / pme_self (:738) /
  pol_penalty (:760); jax.grad of them w.r.t. positions, box, Q_local, Uind_global, tholes, mScales;
  optimize_Uind (:111) -> U, flag, n_cycle; get_forces (:108);
  energy_disp_pme (disp_pme.py:80) for pmax 6/8/10 with gradients;
  generate_pairwise_interaction(TT_damping_qq_c6_kernel) (pairwise.py:45-113) with gradients;
  generate_pme_recip (recip.py:21) stand-alone for Ck_1 (lmax 0,1,2) and Ck_6/8/10.
dE/dpScales and dE/dpol are NaN in the reference on water and are not recorded.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refcases, refrun   # noqa: E402


def main(names):
    ns = refrun.load_reference()
    A, T = refrun.A, refrun.T
    jax = ns.jax
    for name in names:
        t0 = time.time()
        c = refcases.get(name)
        s = c.s
        out = {'positions': s.positions.numpy(), 'box': s.box.numpy(), 'n_pairs': np.int64(c.n_pairs),
               'pairs_checksum': np.int64(int(c.pairs[:c.n_pairs].astype(np.int64).sum()))}
        cov = s.covalent_map.dense()
        pos, box, pairs = A(s.positions), A(s.box), A(c.pairs.astype(np.int64))
        mS, pS, dS = A(s.mScales), A(s.pScales), A(s.dScales)
        full = name in refcases.FULL

        def put(key, val):
            val = T(val)
            out[key] = val.numpy() if isinstance(val, torch.Tensor) else np.asarray(val)

        # ---------------------------------------------------------------- non-polarizable energy_pme
        if name != 'c2':
            f = ns.pme.ADMPPmeForce(box, s.axis_type, s.axis_indices, cov, c.rc, c.ethresh, 2, False)
            if c.kappa is not None:
                f.update_env('kappa', c.kappa)
            put('K', np.array([f.K1, f.K2, f.K3]))
            put('kappa', float(f.kappa))
            Ql = A(s.Q_local if full else c.Q_pert)
            m_in = mS if full else A(c.mScales_pert)
            E, g = jax.value_and_grad(f.get_energy, argnums=(0, 1, 3, 4))(pos, box, pairs, Ql, m_in)
            put('nonpol_E', E); put('nonpol_dpos', g[0]); put('nonpol_dbox', g[1]); put('nonpol_dQ', g[2]); put('nonpol_dmScales', g[3])
            # the three terms, called as energy_pme calls them (pme.py:220-249)
            frames = f.construct_local_frames(pos, box)
            Qg = ns.multipole.rot_local2global(Ql, frames, 2)
            put('nonpol_real', ns.pme.pme_real(pos, box, pairs, Qg, None, None, None, m_in, None, None, cov, f.kappa, 2, False))
            put('nonpol_recip', f.pme_recip(pos, box, Qg))
            put('nonpol_self', ns.pme.pme_self(Qg, f.kappa, 2))
            put('local_frames', frames); put('Q_global', Qg)

        # ---------------------------------------------------------------- polarizable
        if name != 'c1':
            f = ns.pme.ADMPPmeForce(box, s.axis_type, s.axis_indices, cov, c.rc, c.ethresh, 2, True)
            if c.kappa is not None:
                f.update_env('kappa', c.kappa)
            put('K', np.array([f.K1, f.K2, f.K3]))
            put('kappa', float(f.kappa))
            pol, th, Ql = A(s.pol), A(s.tholes), A(s.Q_local)
            if not full:
                # energy_fn at a prescribed U with generic parameters: every coefficient of calc_e_ind matters
                Qp, Up, polp, thp, mp = A(c.Q_pert), A(c.U_pert), A(c.pol_pert), A(c.tholes_pert), A(c.mScales_pert)
                E, g = jax.value_and_grad(f.energy_fn, argnums=(0, 1, 3, 4, 6, 7))(pos, box, pairs, Qp, Up, polp, thp, mp, pS, dS)
                put('pol_E', E); put('pol_dpos', g[0]); put('pol_dbox', g[1]); put('pol_dQ', g[2]); put('pol_dU', g[3])
                put('pol_dtholes', g[4]); put('pol_dmScales', g[5])
                frames = f.construct_local_frames(pos, box)
                Qg = ns.multipole.rot_local2global(Qp, frames, 2)
                Uh = ns.multipole.C1_c2h.dot(Up.T).T
                put('pol_real', ns.pme.pme_real(pos, box, pairs, Qg, Uh, polp, thp, mp, pS, dS, cov, f.kappa, 2, True))
                put('pol_recip', f.pme_recip(pos, box, Qg.at[:, 1:4].add(Uh)))
                put('pol_self', ns.pme.pme_self(Qg.at[:, 1:4].add(Uh), f.kappa, 2))
                put('pol_penalty', ns.pme.pol_penalty(Uh, polp))
            # the SCF and the wrapped energy (water parameters, U_init = the zeros default)
            E, g = jax.value_and_grad(f.get_energy, argnums=(0, 1))(pos, box, pairs, Ql, pol, th, mS, pS, dS)
            put('scf_E', E); put('scf_dpos', g[0]); put('scf_dbox', g[1])
            put('scf_U', f.U_ind); put('scf_n_cycle', np.int64(f.n_cycle)); put('scf_converged', np.bool_(f.lconverg))
            print('  %s: SCF n_cycle %d converged %s  E %.9f' % (name, f.n_cycle, f.lconverg, float(E)), flush=True)

        # ---------------------------------------------------------------- dispersion PME
        if name != 'c2':
            for pmax in ((10,) if full else (6, 8, 10)):
                d = ns.disp_pme.ADMPDispPmeForce(box, cov, c.rc, c.ethresh, pmax)
                if c.kappa is not None:
                    d.update_env('kappa', c.kappa)
                cl = A(s.c_list if full else c.c_list_pert)
                m_in = mS if full else A(c.mScales_pert)
                E, g = jax.value_and_grad(d.get_energy, argnums=(0, 1, 3, 4))(pos, box, pairs, cl, m_in)
                k = 'disp%d_' % pmax
                put(k + 'E', E); put(k + 'dpos', g[0]); put(k + 'dbox', g[1]); put(k + 'dc', g[2]); put(k + 'dmScales', g[3])
                put(k + 'real', ns.disp_pme.disp_pme_real(pos, box, pairs, cl, m_in, cov, d.kappa, pmax))
                put(k + 'self', ns.disp_pme.disp_pme_self(cl, d.kappa, pmax))

        # ---------------------------------------------------------------- TT pair interaction
        if name != 'c2':
            tt = ns.pairwise.generate_pairwise_interaction(ns.pairwise.TT_damping_qq_c6_kernel, cov, {})
            m_in = mS if full else A(c.mScales_pert)
            args = (pos, box, pairs, m_in, A(s.tt_a), A(s.tt_b), A(s.tt_q), A(s.c_list[:, 0].contiguous()))
            E, g = jax.value_and_grad(tt, argnums=(0, 3, 4, 5, 6, 7))(*args)
            put('tt_E', E); put('tt_dpos', g[0]); put('tt_dmScales', g[1]); put('tt_da', g[2]); put('tt_db', g[3])
            put('tt_dq', g[4]); put('tt_dc', g[5])

        # ---------------------------------------------------------------- generate_pme_recip stand-alone
        if not full:
            K = int(out['K'][0])
            kap = float(out['kappa'])
            rng = np.random.default_rng(23)
            Qr = A(rng.normal(0, 0.3, (s.n_atoms, 9)))
            for lmax in (0, 1, 2):
                fn = ns.recip.generate_pme_recip(ns.recip.Ck_1, kap, False, 6, K, K, K, lmax)
                E, g = jax.value_and_grad(fn, argnums=(0, 1, 2))(pos, box, Qr[:, :(lmax + 1) ** 2])
                k = 'recip_l%d_' % lmax
                put(k + 'E', E); put(k + 'dpos', g[0]); put(k + 'dbox', g[1]); put(k + 'dQ', g[2])
            put('recip_Q', Qr)
            for kind, ck in ((6, ns.recip.Ck_6), (8, ns.recip.Ck_8), (10, ns.recip.Ck_10)):
                fn = ns.recip.generate_pme_recip(ck, kap, True, 6, K, K, K, 0)
                E, g = jax.value_and_grad(fn, argnums=(0, 1, 2))(pos, box, Qr[:, :1])
                k = 'recip_c%d_' % kind
                put(k + 'E', E); put(k + 'dpos', g[0]); put(k + 'dbox', g[1]); put(k + 'dQ', g[2])

        path = os.path.join(HERE, 'ref_%s.npz' % name)
        np.savez_compressed(path, **out)
        print('%s: %d arrays -> %s (%.1f s)' % (name, len(out), path, time.time() - t0), flush=True)


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count() or 1)
    main(sys.argv[1:] or list(refcases.SMALL))
