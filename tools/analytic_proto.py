"""NumPy prototype of the analytic formulation the CUDA kernels implement.

NOT product code and NOT the oracle: it is the derivation, written once in NumPy so
that tests/test_analytic_proto.py can check every closed-form adjoint against the
oracle's autograd on the CPU before (and independently of) the CUDA transcription.
The CUDA kernels in admp_b200/csrc follow these formulas line by line.

Formulation (DESIGN.md "Math"):
  * per-site multipoles are Cartesian: q, mu(3), Theta (symmetric traceless, Stone
    convention: Q20 = Tzz, Q21c = 2/sqrt3 Txz, Q21s = 2/sqrt3 Tyz,
    Q22c = (Txx - Tyy)/sqrt3, Q22s = 2/sqrt3 Txy);
  * the pair energy is a sum of rotational invariants of (n, mu, Theta, u) times radial
    functions A0..A9, B1..B6, C2, C3 that are linear combinations of the reference's
    QI-frame coefficients (admp/pme.py:258-334, :379-475);
  * site layout M[:, 0:10] = q, mx, my, mz, Txx, Txy, Txz, Tyy, Tyz, Tzz and the
    gradient G uses the same layout with "one variable per off-diagonal" convention.
"""
import math

import numpy as np

DIEL = 1389.35455846
SQRT3 = math.sqrt(3.0)
SQRT_PI = math.sqrt(math.pi)
THOLE_DEFAULT = 0.3


# --------------------------------------------------------------------------- multipoles
def harm_to_cart(Q):
    """(n,9) harmonic -> (n,10) Cartesian site layout."""
    n = Q.shape[0]
    M = np.zeros((n, 10))
    M[:, 0] = Q[:, 0]
    M[:, 1], M[:, 2], M[:, 3] = Q[:, 2], Q[:, 3], Q[:, 1]
    h = SQRT3 / 2
    M[:, 4] = -0.5 * Q[:, 4] + h * Q[:, 7]   # xx
    M[:, 5] = h * Q[:, 8]                    # xy
    M[:, 6] = h * Q[:, 5]                    # xz
    M[:, 7] = -0.5 * Q[:, 4] - h * Q[:, 7]   # yy
    M[:, 8] = h * Q[:, 6]                    # yz
    M[:, 9] = Q[:, 4]                        # zz
    return M


def cart_grad_to_harm(G):
    """Adjoint of harm_to_cart: (n,10) dE/dM -> (n,9) dE/dQ."""
    n = G.shape[0]
    g = np.zeros((n, 9))
    h = SQRT3 / 2
    g[:, 0] = G[:, 0]
    g[:, 2], g[:, 3], g[:, 1] = G[:, 1], G[:, 2], G[:, 3]
    g[:, 4] = G[:, 9] - 0.5 * G[:, 4] - 0.5 * G[:, 7]
    g[:, 5] = h * G[:, 6]
    g[:, 6] = h * G[:, 8]
    g[:, 7] = h * (G[:, 4] - G[:, 7])
    g[:, 8] = h * G[:, 5]
    return g


def theta_mat(M):
    T = np.empty((M.shape[0], 3, 3))
    T[:, 0, 0], T[:, 0, 1], T[:, 0, 2] = M[:, 4], M[:, 5], M[:, 6]
    T[:, 1, 0], T[:, 1, 1], T[:, 1, 2] = M[:, 5], M[:, 7], M[:, 8]
    T[:, 2, 0], T[:, 2, 1], T[:, 2, 2] = M[:, 6], M[:, 8], M[:, 9]
    return T


def sym_to_layout(Gm):
    """full (n,3,3) matrix derivative (entries independent) -> 6 one-variable comps."""
    return np.stack([Gm[:, 0, 0], Gm[:, 0, 1] + Gm[:, 1, 0], Gm[:, 0, 2] + Gm[:, 2, 0],
                     Gm[:, 1, 1], Gm[:, 1, 2] + Gm[:, 2, 1], Gm[:, 2, 2]], axis=1)


# --------------------------------------------------------------------------- radial functions
def radial_perm(r, m, kappa):
    """A0..A9 and their r-derivatives and m-derivatives.  Returns (A, dA, dAm) each (10, np)."""
    x = kappa * r
    X = 2 * np.exp(-x * x) / SQRT_PI
    ri = [DIEL * r ** (-i) for i in range(7)]
    b1 = -math_erf(x)
    b2 = b1 + x * X
    b3 = b2 + 2 * x**3 * X / 3
    b4 = b3 + 4 * x**5 * X / 15
    db2 = -2 * kappa * x**2 * X
    db3 = -(4 / 3) * kappa * x**4 * X
    db4 = -(8 / 15) * kappa * x**6 * X

    def dxnX(n):          # d/dr (x^n X)
        return kappa * (n * x ** (n - 1) - 2 * x ** (n + 1)) * X

    def dri(i):           # d/dr (DIEL r^-i)
        return -i * ri[i] / r

    # reference coefficients, their r-derivatives and m-derivatives
    cc, dcc, mcc = ri[1] * (m + b2 - x * X), dri(1) * (m + b2 - x * X) + ri[1] * (db2 - dxnX(1)), ri[1]
    cd, dcd, mcd = ri[2] * (m + b2), dri(2) * (m + b2) + ri[2] * db2, ri[2]
    t = 3 * (m + b3) + x**3 * X
    dd0, ddd0, mdd0 = -2 / 3 * ri[3] * t, -2 / 3 * (dri(3) * t + ri[3] * (3 * db3 + dxnX(3))), -2 * ri[3]
    t = m + b3 - (2 / 3) * x**3 * X
    dd1, ddd1, mdd1 = ri[3] * t, dri(3) * t + ri[3] * (db3 - (2 / 3) * dxnX(3)), ri[3]
    cq, dcq, mcq = ri[3] * (m + b3), dri(3) * (m + b3) + ri[3] * db3, ri[3]
    t = 3 * (m + b3) + (4 / 3) * x**5 * X
    dq0, ddq0, mdq0 = ri[4] * t, dri(4) * t + ri[4] * (3 * db3 + (4 / 3) * dxnX(5)), 3 * ri[4]
    dq1, ddq1, mdq1 = -SQRT3 * ri[4] * (m + b3), -SQRT3 * (dri(4) * (m + b3) + ri[4] * db3), -SQRT3 * ri[4]
    t = 6 * (m + b4) + (4 / 45) * (-3 * x**5 + 10 * x**7) * X
    dt = 6 * db4 + (4 / 45) * (-3 * dxnX(5) + 10 * dxnX(7))
    qq0, dqq0, mqq0 = ri[5] * t, dri(5) * t + ri[5] * dt, 6 * ri[5]
    t = 15 * (m + b4) + x**5 * X
    qq1, dqq1, mqq1 = -(4 / 15) * ri[5] * t, -(4 / 15) * (dri(5) * t + ri[5] * (15 * db4 + dxnX(5))), -4 * ri[5]
    t = m + b4 - (4 / 15) * x**5 * X
    qq2, dqq2, mqq2 = ri[5] * t, dri(5) * t + ri[5] * (db4 - (4 / 15) * dxnX(5)), ri[5]

    def combine(cc, cd, dd0, dd1, cq, dq0, dq1, qq0, qq1, qq2):
        c = 2 / SQRT3
        return np.stack([cc, cd, dd0 - dd1, dd1, cq, dq0 - c * dq1, c * dq1,
                         qq0 - (4 / 3) * qq1 + (1 / 3) * qq2, (4 / 3) * (qq1 - qq2), (2 / 3) * qq2])
    A = combine(cc, cd, dd0, dd1, cq, dq0, dq1, qq0, qq1, qq2)
    dA = combine(dcc, dcd, ddd0, ddd1, dcq, ddq0, ddq1, dqq0, dqq1, dqq2)
    mA = combine(*[np.broadcast_to(v, r.shape) for v in (mcc, mcd, mdd0, mdd1, mcq, mdq0, mdq1, mqq0, mqq1, mqq2)])
    return A, dA, mA


def math_erf(x):
    from scipy.special import erf
    return erf(x)


def radial_ind(r, p, th1, th2, pol1, pol2, kappa):
    """B1,B2,B3,B5,B6,C2,C3 (stacked in that order) with r-, p-, a-(thole width) and
    dmp-derivatives.  Follows admp/pme.py:379-475 with the halved perm-induced factors."""
    x = kappa * r
    X = 2 * np.exp(-x * x) / SQRT_PI
    ri = [DIEL * r ** (-i) for i in range(7)]
    b1 = -math_erf(x)
    b2 = b1 + x * X
    b3 = b2 + 2 * x**3 * X / 3
    db2 = -2 * kappa * x**2 * X
    db3 = -(4 / 3) * kappa * x**4 * X

    def dxnX(n):
        return kappa * (n * x ** (n - 1) - 2 * x ** (n + 1)) * X

    def dri(i):
        return -i * ri[i] / r
    # Thole width: piecewise constant in pscale
    uarg = np.minimum((p - 1e-3) / 1e-5, 700.0)
    w0 = 1.0 / (np.exp(uarg) + 1.0)
    a = w0 * THOLE_DEFAULT + (1 - w0) * (th1 + th2)
    da_dth = (1 - w0)
    prod = pol1 * pol2
    dmp_raw = np.where(prod < 1e-48, 0.0, np.maximum(prod, 1e-48) ** (1 / 6))
    trimmed = dmp_raw < 1e-8
    dmp = np.where(trimmed, 1e-8, dmp_raw)
    u_raw = r / dmp
    clipped = u_raw >= 1e8
    u = np.where(clipped, 1e8, u_raw)
    au = a * u
    live = au < 50
    aus = np.where(live, au, 0.0)
    e = np.where(live, np.exp(-aus), 0.0)
    au2, au3, au4 = aus * aus, aus**3, aus**4
    t_c = 1 - e * (1 + aus + 0.5 * au2)
    t_d0 = 1 - e * (1 + aus + 0.5 * au2 + au3 / 4)
    t_q0 = 1 - e * (1 + aus + 0.5 * au2 + au3 / 6 + au4 / 18)
    t_q1 = 1 - e * (1 + aus + 0.5 * au2 + au3 / 6)
    # d/d(au)
    s_c = e * au2 / 2
    s_d0 = e * (au3 - au2) / 4
    s_q0 = e * (au4 - au3) / 18
    s_q1 = e * au3 / 6
    t_d1, s_d1 = t_c, s_c
    # d(au)/dr, /da, /ddmp (zero where u is clipped)
    au_r = np.where(clipped, 0.0, a / dmp)
    au_a = u
    au_d = np.where(clipped, 0.0, -a * r / dmp**2)

    out, d_r, d_p, d_au = [], [], [], []

    def push(val, dr_, dp_, dau_):
        out.append(val), d_r.append(dr_), d_p.append(dp_), d_au.append(dau_)

    # B1 = cud/2
    push(ri[2] * (p * t_c + b2), dri(2) * (p * t_c + b2) + ri[2] * db2, ri[2] * t_c, ri[2] * p * s_c)
    # dud0/2, dud1/2
    t = 3 * (p * t_d0 + b3) + x**3 * X
    h0 = (-2 / 3 * ri[3] * t, -2 / 3 * (dri(3) * t + ri[3] * (3 * db3 + dxnX(3))), -2 * ri[3] * t_d0, -2 * ri[3] * p * s_d0)
    t = p * t_d1 + b3 - (2 / 3) * x**3 * X
    h1 = (ri[3] * t, dri(3) * t + ri[3] * (db3 - (2 / 3) * dxnX(3)), ri[3] * t_d1, ri[3] * p * s_d1)
    push(*[h0[k] - h1[k] for k in range(4)])      # B2
    push(*h1)                                      # B3
    # udq0/2, udq1/2
    t = 3 * (p * t_q0 + b3) + (4 / 3) * x**5 * X
    q0 = (ri[4] * t, dri(4) * t + ri[4] * (3 * db3 + (4 / 3) * dxnX(5)), 3 * ri[4] * t_q0, 3 * ri[4] * p * s_q0)
    t = p * t_q1 + b3
    q1 = (-SQRT3 * ri[4] * t, -SQRT3 * (dri(4) * t + ri[4] * db3), -SQRT3 * ri[4] * t_q1, -SQRT3 * ri[4] * p * s_q1)
    c = 2 / SQRT3
    push(*[q0[k] - c * q1[k] for k in range(4)])   # B5
    push(*[c * q1[k] for k in range(4)])           # B6
    # udud0, udud1 (uscales = 1)
    t = 3 * (t_d0 + b3) + x**3 * X
    u0 = (-2 / 3 * ri[3] * t, -2 / 3 * (dri(3) * t + ri[3] * (3 * db3 + dxnX(3))), 0 * r, -2 * ri[3] * s_d0)
    t = t_d1 + b3 - (2 / 3) * x**3 * X
    u1 = (ri[3] * t, dri(3) * t + ri[3] * (db3 - (2 / 3) * dxnX(3)), 0 * r, ri[3] * s_d1)
    push(*[u0[k] - u1[k] for k in range(4)])       # C2
    push(*u1)                                      # C3
    B, dB_au = np.stack(out), np.stack(d_au)
    dB_r = np.stack(d_r) + dB_au * au_r
    extra = dict(dB_p=np.stack(d_p), dB_a=dB_au * au_a, dB_dmp=dB_au * au_d, da_dth=da_dth,
                 trimmed=trimmed, dmp=dmp)
    return B, dB_r, extra


# --------------------------------------------------------------------------- pair kernel
def pair_real(pos, box_l, pairs, M, U, pol, tholes, mS, pS, scale_idx, kappa, lpol):
    """Real-space energy and all adjoints (orthorhombic box lengths box_l (3,)).

    Returns dict(E, dpos (n,3), G (n,10), F (n,3) = dE/dU, dbox (3,3), dmS (5), dpS(5),
    dthole (n), dpol (n)).
    """
    n_atoms = pos.shape[0]
    i, j = pairs[:, 0], pairs[:, 1]
    d = pos[i] - pos[j]
    shift = np.floor(d / box_l + 0.5)
    d = d - shift * box_l
    r = np.linalg.norm(d, axis=1)
    n = d / r[:, None]
    m = mS[scale_idx]
    A, dA, mA = radial_perm(r, m, kappa)
    qI, qJ = M[i, 0], M[j, 0]
    muI, muJ = M[i, 1:4], M[j, 1:4]
    TI, TJ = theta_mat(M[i]), theta_mat(M[j])
    dI, dJ = np.sum(muI * n, 1), np.sum(muJ * n, 1)
    vI, vJ = np.einsum('pab,pb->pa', TI, n), np.einsum('pab,pb->pa', TJ, n)
    tI, tJ = np.sum(vI * n, 1), np.sum(vJ * n, 1)
    mm = np.sum(muI * muJ, 1)
    gJI, gIJ = np.sum(muJ * vI, 1), np.sum(muI * vJ, 1)
    vv = np.sum(vI * vJ, 1)
    TT = np.einsum('pab,pab->p', TI, TJ)
    inv = np.stack([qI * qJ, qI * dJ - dI * qJ, dI * dJ, mm, tI * qJ + qI * tJ, tI * dJ - dI * tJ,
                    gJI - gIJ, tI * tJ, vv, TT])
    E = np.sum(A * inv, 0)
    dEdr = np.sum(dA * inv, 0)
    dmS = np.zeros(5)
    np.add.at(dmS, scale_idx % 5, np.sum(mA * inv, 0))
    # partial derivatives w.r.t. the scalar invariants' building blocks
    e_dI = -A[1] * qJ + A[2] * dJ - A[5] * tJ
    e_dJ = A[1] * qI + A[2] * dI + A[5] * tI
    e_tI = A[4] * qJ + A[5] * dJ + A[7] * tJ
    e_tJ = A[4] * qI - A[5] * dI + A[7] * tI
    g_qI = A[0] * qJ + A[1] * dJ + A[4] * tJ
    g_qJ = A[0] * qI - A[1] * dI + A[4] * tI
    e_pI = e_pJ = 0
    if lpol:
        B, dB, ex = radial_ind(r, pS[scale_idx], tholes[i], tholes[j], pol[i], pol[j], kappa)
        uI, uJ = U[i], U[j]
        pI, pJ = np.sum(uI * n, 1), np.sum(uJ * n, 1)
        inv2 = np.stack([qI * pJ - pI * qJ, pI * dJ + pJ * dI, np.sum(uI * muJ, 1) + np.sum(uJ * muI, 1),
                         tI * pJ - pI * tJ, np.sum(uJ * vI, 1) - np.sum(uI * vJ, 1), pI * pJ, np.sum(uI * uJ, 1)])
        E = E + np.sum(B * inv2, 0)
        dEdr = dEdr + np.sum(dB * inv2, 0)
        B1, B2, B3, B5, B6, C2, C3 = B
        e_dI = e_dI + B2 * pJ
        e_dJ = e_dJ + B2 * pI
        e_tI = e_tI + B5 * pJ
        e_tJ = e_tJ - B5 * pI
        g_qI = g_qI + B1 * pJ
        g_qJ = g_qJ - B1 * pI
        e_pI = -B1 * qJ + B2 * dJ - B5 * tJ + C2 * pJ
        e_pJ = B1 * qI + B2 * dI + B5 * tI + C2 * pI
    g_muI = e_dI[:, None] * n + A[3][:, None] * muJ - A[6][:, None] * vJ
    g_muJ = e_dJ[:, None] * n + A[3][:, None] * muI + A[6][:, None] * vI
    nn = n[:, :, None] * n[:, None, :]
    G_TI = e_tI[:, None, None] * nn + A[6][:, None, None] * muJ[:, :, None] * n[:, None, :] \
        + A[8][:, None, None] * vJ[:, :, None] * n[:, None, :] + A[9][:, None, None] * TJ
    G_TJ = e_tJ[:, None, None] * nn - A[6][:, None, None] * muI[:, :, None] * n[:, None, :] \
        + A[8][:, None, None] * vI[:, :, None] * n[:, None, :] + A[9][:, None, None] * TI
    gn = e_dI[:, None] * muI + e_dJ[:, None] * muJ + 2 * e_tI[:, None] * vI + 2 * e_tJ[:, None] * vJ \
        + A[6][:, None] * (np.einsum('pab,pb->pa', TI, muJ) - np.einsum('pab,pb->pa', TJ, muI)) \
        + A[8][:, None] * (np.einsum('pab,pb->pa', TI, vJ) + np.einsum('pab,pb->pa', TJ, vI))
    F = np.zeros((n_atoms, 3))
    dpS = np.zeros(5)
    dth = np.zeros(n_atoms)
    dpol = np.zeros(n_atoms)
    if lpol:
        g_muI = g_muI + B3[:, None] * uJ
        g_muJ = g_muJ + B3[:, None] * uI
        G_TI = G_TI + B6[:, None, None] * uJ[:, :, None] * n[:, None, :]
        G_TJ = G_TJ - B6[:, None, None] * uI[:, :, None] * n[:, None, :]
        g_uI = e_pI[:, None] * n + B3[:, None] * muJ - B6[:, None] * vJ + C3[:, None] * uJ
        g_uJ = e_pJ[:, None] * n + B3[:, None] * muI + B6[:, None] * vI + C3[:, None] * uI
        gn = gn + e_pI[:, None] * uI + e_pJ[:, None] * uJ \
            + B6[:, None] * (np.einsum('pab,pb->pa', TI, uJ) - np.einsum('pab,pb->pa', TJ, uI))
        np.add.at(F, i, g_uI)
        np.add.at(F, j, g_uJ)
        np.add.at(dpS, scale_idx % 5, np.sum(ex['dB_p'] * inv2, 0))
        e_a = np.sum(ex['dB_a'] * inv2, 0) * ex['da_dth']
        np.add.at(dth, i, e_a)
        np.add.at(dth, j, e_a)
        e_dmp = np.where(ex['trimmed'], 0.0, np.sum(ex['dB_dmp'] * inv2, 0))
        with np.errstate(divide='ignore', invalid='ignore'):
            np.add.at(dpol, i, np.where(ex['trimmed'], 0.0, e_dmp * ex['dmp'] / (6 * pol[i])))
            np.add.at(dpol, j, np.where(ex['trimmed'], 0.0, e_dmp * ex['dmp'] / (6 * pol[j])))
    fvec = dEdr[:, None] * n + (gn - np.sum(gn * n, 1)[:, None] * n) / r[:, None]
    dpos = np.zeros((n_atoms, 3))
    np.add.at(dpos, i, fvec)
    np.add.at(dpos, j, -fvec)
    G = np.zeros((n_atoms, 10))
    np.add.at(G, i, np.concatenate([g_qI[:, None], g_muI, sym_to_layout(G_TI)], 1))
    np.add.at(G, j, np.concatenate([g_qJ[:, None], g_muJ, sym_to_layout(G_TJ)], 1))
    dbox = -np.einsum('pa,pb->ab', shift, fvec)
    return dict(E=E.sum(), dpos=dpos, G=G, F=F, dbox=dbox, dmS=dmS, dpS=dpS, dthole=dth, dpol=dpol)


# --------------------------------------------------------------------------- B-splines
def bspline6_all(f):
    """w[p][k] = d^p/du^p M6(u) at u = f + k, k = 0..5, p = 0..3.  f (n,) in [0,1)."""
    n = f.shape[0]

    def step(prev, order):
        cur = np.zeros((order, n))
        for k in range(order):
            a = prev[k] if k < order - 1 else 0.0
            b = prev[k - 1] if k >= 1 else 0.0
            cur[k] = ((f + k) * a + (order - f - k) * b) / (order - 1)
        return cur
    a2 = np.stack([f, 1 - f])
    a3 = step(a2, 3)
    a4 = step(a3, 4)
    a5 = step(a4, 5)
    a6 = step(a5, 6)

    def pad(a):
        z = np.zeros((9, n))
        z[3:3 + a.shape[0]] = a
        return z          # index k+3
    p5, p4, p3 = pad(a5), pad(a4), pad(a3)
    w = np.zeros((4, 6, n))
    for k in range(6):
        w[0, k] = a6[k]
        w[1, k] = p5[k + 3] - p5[k + 2]
        w[2, k] = p4[k + 3] - 2 * p4[k + 2] + p4[k + 1]
        w[3, k] = p3[k + 3] - 3 * p3[k + 2] + 3 * p3[k + 1] - p3[k]
    return w


# --------------------------------------------------------------------------- reciprocal space
def recip_setup(pos, box, K):
    K = np.asarray(K, dtype=np.float64)
    inv = np.linalg.inv(box)
    Nstar = (K[None, :] * inv).T
    x = pos @ Nstar.T
    m0 = np.ceil(x)
    f = m0 - x
    w = [bspline6_all(f[:, d]) for d in range(3)]      # w[d][p,k,atom]
    return Nstar, inv, m0.astype(np.int64), w


def frac_multipoles(M, Nstar):
    """fractional coefficient table F[atom, (p1,p2,p3)] for the 10 derivative orders <= 2."""
    mu = M[:, 1:4]
    T = theta_mat(M)
    muf = -mu @ Nstar.T                                   # (n,3): -sum_c Nstar[d][c] mu_c
    Tf = np.einsum('da,eb,nab->nde', Nstar, Nstar, T) / 3.0
    return M[:, 0], muf, Tf


def spread(pos, box, K, M):
    Nstar, inv, m0, w = recip_setup(pos, box, K)
    q, muf, Tf = frac_multipoles(M, Nstar)
    mesh = np.zeros(tuple(int(k) for k in K))
    n = pos.shape[0]
    ks = np.arange(6)
    for a in range(n):
        w0 = [w[d][0, :, a] for d in range(3)]
        w1 = [w[d][1, :, a] for d in range(3)]
        w2 = [w[d][2, :, a] for d in range(3)]
        val = q[a] * np.einsum('i,j,k->ijk', w0[0], w0[1], w0[2])
        val += muf[a, 0] * np.einsum('i,j,k->ijk', w1[0], w0[1], w0[2])
        val += muf[a, 1] * np.einsum('i,j,k->ijk', w0[0], w1[1], w0[2])
        val += muf[a, 2] * np.einsum('i,j,k->ijk', w0[0], w0[1], w1[2])
        val += Tf[a, 0, 0] * np.einsum('i,j,k->ijk', w2[0], w0[1], w0[2])
        val += Tf[a, 1, 1] * np.einsum('i,j,k->ijk', w0[0], w2[1], w0[2])
        val += Tf[a, 2, 2] * np.einsum('i,j,k->ijk', w0[0], w0[1], w2[2])
        val += 2 * Tf[a, 0, 1] * np.einsum('i,j,k->ijk', w1[0], w1[1], w0[2])
        val += 2 * Tf[a, 0, 2] * np.einsum('i,j,k->ijk', w1[0], w0[1], w1[2])
        val += 2 * Tf[a, 1, 2] * np.einsum('i,j,k->ijk', w0[0], w1[1], w1[2])
        ix = (m0[a, 0] - 3 + ks) % int(K[0])
        iy = (m0[a, 1] - 3 + ks) % int(K[1])
        iz = (m0[a, 2] - 3 + ks) % int(K[2])
        np.add.at(mesh, (ix[:, None, None], iy[None, :, None], iz[None, None, :]), val)
    return mesh


def kint(n):
    k = np.arange(n)
    return np.where(k <= (n - 1) // 2 if n % 2 else k < n // 2, k, k - n)


def theta_inv2(n):
    """1/theta_k^2 per dimension."""
    from scipy.interpolate import BSpline  # noqa: F401  (not used; explicit formula below)
    w = bspline6_all(np.array([0.0]))[0, :, 0]        # M6(k), k = 0..5 ; M6(m+3) for m=-2..2 -> k=1..5
    m = np.arange(-2, 3)
    k = kint(n)
    th = np.sum(w[1:6][:, None] * np.cos(2 * np.pi * m[:, None] * k[None, :] / n), 0)
    return 1.0 / th**2


def ck_and_deriv(ksq, kappa, V, kind):
    """C_k and dC_k/d(k^2)."""
    from scipy.special import erfc
    if kind == 1:
        with np.errstate(divide='ignore', invalid='ignore'):
            C = 2 * np.pi / V / ksq * np.exp(-ksq / 4 / kappa**2)
            dC = -C * (1 / ksq + 1 / (4 * kappa**2))
        C = np.where(ksq == 0, 0.0, C)
        dC = np.where(ksq == 0, 0.0, dC)
        return C, dC
    x2 = ksq / 4 / kappa**2
    x = np.sqrt(x2)
    e = np.exp(-x2)
    ec = SQRT_PI * erfc(x)
    if kind == 6:
        f = (1 - 2 * x2) * e + 2 * x2 * x * ec
        df = (-6 * e + 6 * x * ec)
        pref = SQRT_PI * np.pi / 2 / V * kappa**3 / 3
    elif kind == 8:
        f = (3 - 2 * x2 + 4 * x2**2) * e - 4 * x2**2 * x * ec
        df = (e * (-10 + 20 * x2) - 20 * x2 * x * ec)
        pref = SQRT_PI * np.pi / 2 / V * kappa**5 / 45
    else:
        f = (15 - 6 * x2 + 4 * x2**2 - 8 * x2**3) * e + 8 * x2**3 * x * ec
        df = (e * (-42 + 28 * x2 - 56 * x2**2) + 56 * x2**2 * x * ec)
        pref = SQRT_PI * np.pi / 2 / V * kappa**7 / 1260
    return pref * f, pref * df / (8 * kappa**2)


def recip_all(pos, box, K, M, kappa, kind=1):
    """Reciprocal energy with all adjoints: E, dpos, G (n,10), dbox (3,3)."""
    K = [int(k) for k in K]
    Nstar, inv, m0, w = recip_setup(pos, box, K)
    mesh = spread(pos, box, K, M)
    S = np.fft.rfftn(mesh)
    k1, k2, k3 = kint(K[0]), kint(K[1]), np.arange(K[2] // 2 + 1)
    kv = 2 * np.pi * (k1[:, None, None, None] * inv[0][None, None, None, :]
                      + k2[None, :, None, None] * inv[1][None, None, None, :]
                      + k3[None, None, :, None] * inv[2][None, None, None, :])
    ksq = np.sum(kv**2, -1)
    V = np.linalg.det(box)
    C, dC = ck_and_deriv(ksq, kappa, V, kind)
    scale = DIEL if kind == 1 else 1.0
    th = theta_inv2(K[0])[:, None, None] * theta_inv2(K[1])[None, :, None] * theta_inv2(K[2])[None, None, :K[2] // 2 + 1]
    wgt = np.full(K[2] // 2 + 1, 2.0)
    wgt[0] = 1.0
    if K[2] % 2 == 0:
        wgt[-1] = 1.0
    S2 = (S.real**2 + S.imag**2) * th * wgt[None, None, :]
    E = scale * np.sum(C * S2)
    # k_a k_c summed over the full spectrum: a weight-2 half-spectrum point stands for
    # itself and its Hermitian partner, whose k is -k except in a dimension where the
    # point sits on the (even-N) Nyquist index, which aliases onto itself.
    nyq = [np.arange(K[d]) == (K[d] // 2 if K[d] % 2 == 0 else -1) for d in range(2)]
    sgn = np.stack(np.broadcast_arrays(np.where(nyq[0], 1.0, -1.0)[:, None, None],
                                       np.where(nyq[1], 1.0, -1.0)[None, :, None],
                                       -np.ones(K[2] // 2 + 1)[None, None, :]), -1)
    kp = kv * sgn
    w1 = (wgt == 1.0)[None, None, :]
    base = scale * dC * (S.real**2 + S.imag**2) * th
    Tk = np.einsum('xyz,xyza,xyzc->ac', base, kv, kv) + np.einsum('xyz,xyza,xyzc->ac', np.where(w1, 0.0, base), kp, kp)
    dbox_k = -2 * Tk @ inv.T - E * inv.T
    phi = 2 * scale * np.fft.irfftn(C * th * S, s=K, axes=(0, 1, 2)) * (K[0] * K[1] * K[2])
    # gather fractional derivatives up to order 3
    n = pos.shape[0]
    q, muf, Tf = frac_multipoles(M, Nstar)
    ks = np.arange(6)
    G = np.zeros((n, 10))
    dpos = np.zeros((n, 3))
    W = np.zeros((3, 3))                       # dE/dNstar[d][c]
    T = theta_mat(M)
    for a in range(n):
        ix = (m0[a, 0] - 3 + ks) % K[0]
        iy = (m0[a, 1] - 3 + ks) % K[1]
        iz = (m0[a, 2] - 3 + ks) % K[2]
        blk = phi[ix[:, None, None], iy[None, :, None], iz[None, None, :]]
        P = np.einsum('ijk,pi,qj,rk->pqr', blk, w[0][:, :, a], w[1][:, :, a], w[2][:, :, a])   # P[p1,p2,p3]
        e = np.eye(3, dtype=int)

        def Ph(*ds):
            idx = np.zeros(3, dtype=int)
            for d_ in ds:
                idx += e[d_]
            return P[idx[0], idx[1], idx[2]]
        ph1 = np.array([Ph(d) for d in range(3)])
        ph2 = np.array([[Ph(d, e_) for e_ in range(3)] for d in range(3)])
        ph3 = np.array([[[Ph(d, e_, f_) for f_ in range(3)] for e_ in range(3)] for d in range(3)])
        G[a, 0] = P[0, 0, 0]
        G[a, 1:4] = -Nstar.T @ ph1
        Gm = Nstar.T @ ph2 @ Nstar / 3.0
        G[a, 4:] = [Gm[0, 0], 2 * Gm[0, 1], 2 * Gm[0, 2], Gm[1, 1], 2 * Gm[1, 2], Gm[2, 2]]
        dEdu = q[a] * ph1 + ph2 @ muf[a] + np.einsum('def,ef->d', ph3, Tf[a])
        dpos[a] = -Nstar.T @ dEdu
        W += np.outer(dEdu, -pos[a]) + np.outer(ph1, -M[a, 1:4]) + 2 * ph2 @ (Nstar @ T[a]) / 3.0
    Mm = W.T @ Nstar                          # M[c][b] = sum_d W[d][c] Nstar[d][b]
    dbox = -(inv.T @ Mm) + dbox_k
    return dict(E=E, dpos=dpos, G=G, dbox=dbox, phi=phi, mesh=mesh)


# --------------------------------------------------------------------------- self / penalty
def self_terms(M, U, pol, kappa, lpol):
    f0 = kappa / SQRT_PI
    f1 = f0 * (2 * kappa**2) / 3
    f2 = f0 * (2 * kappa**2) ** 2 / 15
    mu = M[:, 1:4] + (U if lpol else 0)
    G = np.zeros_like(M)
    E = -DIEL * (f0 * np.sum(M[:, 0] ** 2) + f1 * np.sum(mu**2)
                 + f2 * (2 / 3) * np.sum(M[:, [4, 7, 9]] ** 2 + 2 * M[:, [5, 6, 8]] ** 2))
    G[:, 0] = -2 * DIEL * f0 * M[:, 0]
    G[:, 1:4] = -2 * DIEL * f1 * mu
    G[:, [4, 7, 9]] = -DIEL * f2 * (4 / 3) * M[:, [4, 7, 9]]
    G[:, [5, 6, 8]] = -DIEL * f2 * (8 / 3) * M[:, [5, 6, 8]]
    F = np.zeros_like(mu)
    dpol = np.zeros(M.shape[0])
    if lpol:
        pt = np.maximum(pol, 1e-8)
        E += DIEL * np.sum(0.5 / pt[:, None] * U**2)
        F = G[:, 1:4] + DIEL * U / pt[:, None]
        dpol = np.where(pol >= 1e-8, -DIEL * 0.5 * np.sum(U**2, 1) / pt**2, 0.0)
    return dict(E=E, G=G, F=F, dpol=dpol)


# --------------------------------------------------------------------------- frames
ZTHENX, BISECTOR, ZBISECT, THREEFOLD, ZONLY, NOAXIS = 0, 1, 2, 3, 4, 5


def _unit_fwd(v):
    nv = np.linalg.norm(v)
    return v / nv, nv


def _unit_bwd(g, u, nv):
    return (g - u * np.dot(u, g)) / nv


def frames_fwd_bwd(pos, box_l, axis_type, axis_idx, Q_local, G=None):
    """Forward: frames R (n,3,3) and global Cartesian multipoles M (n,10).
    Backward (if G (n,10) = dE/dM given): dE/dQ_local (n,9), dpos (n,3), dbox (3,3)."""
    n = pos.shape[0]
    Ml = harm_to_cart(Q_local)
    R = np.zeros((n, 3, 3))
    M = np.zeros((n, 10))
    dQ = np.zeros((n, 9))
    dpos = np.zeros((n, 3))
    dbox = np.zeros((3, 3))
    GMl = np.zeros((n, 10))

    def disp(a, b):
        d = pos[b] - pos[a]
        s = np.floor(d / box_l + 0.5)
        return d - s * box_l, s
    for a in range(n):
        t = axis_type[a]
        if t == NOAXIS:
            R[a] = np.eye(3)
            M[a] = Ml[a]
            if G is not None:
                GMl[a] = G[a]
            continue
        dz, sz = disp(a, axis_idx[a, 0])
        vz0, nz = _unit_fwd(dz)
        if t == ZONLY:
            xz0 = np.round(abs(vz0[0]))
            vx0 = np.array([1 - xz0, xz0, 0.0])
        else:
            dx, sx = disp(a, axis_idx[a, 1])
            vx0, nx = _unit_fwd(dx)
        vz, vx = vz0, vx0
        if t == BISECTOR:
            vz, nzb = _unit_fwd(vz0 + vx0)
        if t in (ZBISECT, THREEFOLD):
            dy, sy = disp(a, axis_idx[a, 2])
            vy0, ny = _unit_fwd(dy)
        if t == ZBISECT:
            vx, nxb = _unit_fwd(vx0 + vy0)
        if t == THREEFOLD:
            vz, nzb = _unit_fwd(vz0 + vx0 + vy0)
        s = np.dot(vx, vz)
        wv = vx - vz * s
        vxp, nw = _unit_fwd(wv)
        vy = np.cross(vz, vxp)
        Ra = np.stack([vxp, vy, vz])
        R[a] = Ra
        mul = Ml[a, 1:4]
        Tl = theta_mat(Ml[a:a + 1])[0]
        M[a, 0] = Ml[a, 0]
        M[a, 1:4] = Ra.T @ mul
        Tg = Ra.T @ Tl @ Ra
        M[a, 4:] = [Tg[0, 0], Tg[0, 1], Tg[0, 2], Tg[1, 1], Tg[1, 2], Tg[2, 2]]
        if G is None:
            continue
        gmu = G[a, 1:4]
        Gs = np.array([[G[a, 4], G[a, 5] / 2, G[a, 6] / 2],
                       [G[a, 5] / 2, G[a, 7], G[a, 8] / 2],
                       [G[a, 6] / 2, G[a, 8] / 2, G[a, 9]]])
        GMl[a, 0] = G[a, 0]
        GMl[a, 1:4] = Ra @ gmu
        GTl = Ra @ Gs @ Ra.T
        GMl[a, 4:] = [GTl[0, 0], 2 * GTl[0, 1], 2 * GTl[0, 2], GTl[1, 1], 2 * GTl[1, 2], GTl[2, 2]]
        gR = np.outer(mul, gmu) + 2 * Tl @ Ra @ Gs        # gR[i][a]
        gx, gy, gz = gR[0].copy(), gR[1].copy(), gR[2].copy()
        # vy = vz x vxp
        gz = gz + np.cross(vxp, gy)
        gxp = gx + np.cross(gy, vz)
        gw = _unit_bwd(gxp, vxp, nw)
        gvx = gw - vz * np.dot(vz, gw)
        gz = gz - s * gw - np.dot(gw, vz) * vx
        g_vz0, g_vx0, g_vy0 = gz, gvx, np.zeros(3)
        if t == BISECTOR:
            gsum = _unit_bwd(gz, vz, nzb)
            g_vz0, g_vx0 = gsum, gvx + gsum
        if t == ZBISECT:
            gsum = _unit_bwd(gvx, vx, nxb)
            g_vx0, g_vy0 = gsum, gsum
        if t == THREEFOLD:
            gsum = _unit_bwd(gz, vz, nzb)
            g_vz0, g_vx0, g_vy0 = gsum, gvx + gsum, gsum
        gd = _unit_bwd(g_vz0, vz0, nz)
        dpos[axis_idx[a, 0]] += gd
        dpos[a] -= gd
        dbox -= np.outer(sz, gd)
        if t != ZONLY:
            gd = _unit_bwd(g_vx0, vx0, nx)
            dpos[axis_idx[a, 1]] += gd
            dpos[a] -= gd
            dbox -= np.outer(sx, gd)
        if t in (ZBISECT, THREEFOLD):
            gd = _unit_bwd(g_vy0, vy0, ny)
            dpos[axis_idx[a, 2]] += gd
            dpos[a] -= gd
            dbox -= np.outer(sy, gd)
    if G is not None:
        dQ = cart_grad_to_harm(GMl)
    return dict(R=R, M=M, dQ=dQ, dpos=dpos, dbox=dbox)
