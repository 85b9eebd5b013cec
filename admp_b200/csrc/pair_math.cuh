// Per-pair arithmetic shared by the flat pair kernel (pair.cu) and the cluster pair kernel (pair_cluster.cu):
// radial functions of admp/pme.py:258-475 in rotational-invariant form (see pair.cu), the packed per-atom record,
// and the warp reduce-scatter used for the j-side gradients.
#pragma once
#include "kernels.h"

namespace admp {


template <typename T> struct Radial {
    T x, X, ri[6];   // ri[i] = DIEL r^-i
    T b2, b3, b4, db2, db3, db4, kappa, r;
    __device__ __forceinline__ T dxnX(int n, T xnm1, T xnp1) const { return kappa * ((T)n * xnm1 - 2 * xnp1) * X; }
};

template <typename T>
__device__ __forceinline__ void radial_setup(T r, T rinv, T kappa, Radial<T>& R) {
    R.r = r; R.kappa = kappa;
    R.ri[0] = (T)ADMP_DIEL;
#pragma unroll
    for (int i = 1; i < 6; ++i) R.ri[i] = R.ri[i - 1] * rinv;
    const T x = kappa * r, x2 = x * x;
    R.x = x;
    R.X = (T)(2 / ADMP_SQRT_PI) * exp(-x2);
    const T b1 = -erf(x);
    const T x3 = x2 * x, x5 = x3 * x2;
    R.b2 = b1 + x * R.X;
    R.b3 = R.b2 + (T)(2.0 / 3) * x3 * R.X;
    R.b4 = R.b3 + (T)(4.0 / 15) * x5 * R.X;
    R.db2 = -2 * kappa * x2 * R.X;
    R.db3 = (T)(-4.0 / 3) * kappa * x2 * x2 * R.X;
    R.db4 = (T)(-8.0 / 15) * kappa * x3 * x3 * R.X;
}

// A[0..9] (+ dA/dr, dA/dm) for scale m. With thole factors (t*, s* = dt/d(au)) this same routine
// yields the halved perm-induced coefficients: pass m*t_x per rank via the "eff" arguments.
template <typename T, bool DERIV>
__device__ __forceinline__ void perm_coeffs(const Radial<T>& R, T m, T (&A)[10], T (&dA)[10], T (&mA)[10]) {
    const T x = R.x, x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x, x6 = x3 * x3, x7 = x6 * x, x8 = x4 * x4, X = R.X;
    const T c23 = (T)(2 / ADMP_SQRT3), s3 = (T)ADMP_SQRT3;
    const T* ri = R.ri;
    const T rinv = ri[1] * (T)(1 / ADMP_DIEL);
    // reference coefficients
    const T tcc = m + R.b2 - x * X;
    const T cc = ri[1] * tcc;
    const T cd = ri[2] * (m + R.b2);
    const T tdd0 = 3 * (m + R.b3) + x3 * X;
    const T dd0 = (T)(-2.0 / 3) * ri[3] * tdd0;
    const T tdd1 = m + R.b3 - (T)(2.0 / 3) * x3 * X;
    const T dd1 = ri[3] * tdd1;
    const T cq = ri[3] * (m + R.b3);
    const T tdq0 = 3 * (m + R.b3) + (T)(4.0 / 3) * x5 * X;
    const T dq0 = ri[4] * tdq0;
    const T dq1 = -s3 * ri[4] * (m + R.b3);
    const T tqq0 = 6 * (m + R.b4) + (T)(4.0 / 45) * (-3 * x5 + 10 * x7) * X;
    const T qq0 = ri[5] * tqq0;
    const T tqq1 = 15 * (m + R.b4) + x5 * X;
    const T qq1 = (T)(-4.0 / 15) * ri[5] * tqq1;
    const T tqq2 = m + R.b4 - (T)(4.0 / 15) * x5 * X;
    const T qq2 = ri[5] * tqq2;
    A[0] = cc; A[1] = cd; A[2] = dd0 - dd1; A[3] = dd1; A[4] = cq; A[5] = dq0 - c23 * dq1; A[6] = c23 * dq1;
    A[7] = qq0 - (T)(4.0 / 3) * qq1 + (T)(1.0 / 3) * qq2; A[8] = (T)(4.0 / 3) * (qq1 - qq2); A[9] = (T)(2.0 / 3) * qq2;
    if (DERIV) {
        const T d1 = R.dxnX(1, (T)1, x2), d3 = R.dxnX(3, x2, x4), d5 = R.dxnX(5, x4, x6), d7 = R.dxnX(7, x6, x8);
        const T dcc = -rinv * cc + ri[1] * (R.db2 - d1);
        const T dcd = -2 * rinv * cd + ri[2] * R.db2;
        const T ddd0 = -3 * rinv * dd0 + (T)(-2.0 / 3) * ri[3] * (3 * R.db3 + d3);
        const T ddd1 = -3 * rinv * dd1 + ri[3] * (R.db3 - (T)(2.0 / 3) * d3);
        const T dcq = -3 * rinv * cq + ri[3] * R.db3;
        const T ddq0 = -4 * rinv * dq0 + ri[4] * (3 * R.db3 + (T)(4.0 / 3) * d5);
        const T ddq1 = -4 * rinv * dq1 - s3 * ri[4] * R.db3;
        const T dqq0 = -5 * rinv * qq0 + ri[5] * (6 * R.db4 + (T)(4.0 / 45) * (-3 * d5 + 10 * d7));
        const T dqq1 = -5 * rinv * qq1 + (T)(-4.0 / 15) * ri[5] * (15 * R.db4 + d5);
        const T dqq2 = -5 * rinv * qq2 + ri[5] * (R.db4 - (T)(4.0 / 15) * d5);
        dA[0] = dcc; dA[1] = dcd; dA[2] = ddd0 - ddd1; dA[3] = ddd1; dA[4] = dcq; dA[5] = ddq0 - c23 * ddq1; dA[6] = c23 * ddq1;
        dA[7] = dqq0 - (T)(4.0 / 3) * dqq1 + (T)(1.0 / 3) * dqq2; dA[8] = (T)(4.0 / 3) * (dqq1 - dqq2); dA[9] = (T)(2.0 / 3) * dqq2;
        // d/dm of (cc, cd, dd0, dd1, cq, dq0, dq1, qq0, qq1, qq2) = (r1, r2, -2r3, r3, r3, 3r4, -s3 r4, 6r5, -4r5, r5)
        mA[0] = ri[1]; mA[1] = ri[2]; mA[2] = -3 * ri[3]; mA[3] = ri[3]; mA[4] = ri[3]; mA[5] = 3 * ri[4] + c23 * s3 * ri[4];
        mA[6] = -c23 * s3 * ri[4]; mA[7] = 6 * ri[5] + (T)(16.0 / 3) * ri[5] + (T)(1.0 / 3) * ri[5];
        mA[8] = (T)(4.0 / 3) * (-5 * ri[5]); mA[9] = (T)(2.0 / 3) * ri[5];
    }
}

// B[0..6] = B1, B2, B3, B5, B6, C2, C3 (admp/pme.py:379-475, halved perm-induced factors)
template <typename T> struct IndCoef {
    T B[7], dB[7];        // value, total d/dr
    T pB[7], aB[7];       // d/dpscale, d/d(au)  (for parameter gradients)
    T au_a, au_d, da_dth; // d(au)/da, d(au)/d(dmp), da/dthole
    bool trimmed;
    T dmp;
};

// Fermi switch of the Thole width (admp/pme.py:337-348,411): depends on the pair's pscale only, i.e. on its scale
// index - the kernels tabulate it once per launch instead of one exp + one division per pair.
template <typename T> __device__ __forceinline__ T thole_switch_w0(T p) {
    T uarg = (p - (T)1e-3) * (T)1e5;
    uarg = uarg > (T)80 ? (T)80 : uarg;
    return (T)1 / (exp(uarg) + (T)1);
}

// SIXTH: pol1 / pol2 are pol^(1/6) per atom (packed record slot 18; 0 for pol = 0), so that the pair's
// dmp = (pol1 pol2)^(1/6) is one product instead of a double-precision pow per pair.
template <typename T, bool DERIV, bool SIXTH = false>
__device__ __forceinline__ void ind_coeffs(const Radial<T>& R, T p, T w0, T th1, T th2, T pol1, T pol2, IndCoef<T>& C) {
    const T x = R.x, x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x, x6 = x3 * x3, X = R.X;
    const T c23 = (T)(2 / ADMP_SQRT3), s3 = (T)ADMP_SQRT3;
    const T* ri = R.ri;
    const T rinv = ri[1] * (T)(1 / ADMP_DIEL);
    // Thole width: Fermi switch of admp/pme.py:337-348,411, piecewise constant in pscale (A7)
    const T a = w0 * (T)ADMP_THOLE_DEFAULT + ((T)1 - w0) * (th1 + th2);
    C.da_dth = (T)1 - w0;
    // dmp = trim_val_0((pol1 pol2)^(1/6)), u = trim_val_infty(r/dmp)   (pme.py:413-414,732-735)
    const double prod = (double)pol1 * (double)pol2;
    double dmpd = SIXTH ? prod : (prod < 1e-48 ? 0.0 : pow(prod, 1.0 / 6.0));
    C.trimmed = dmpd < 1e-8;
    if (C.trimmed) dmpd = 1e-8;
    const T dmp = (T)dmpd;
    C.dmp = dmp;
    const double dinv = 1.0 / dmpd;
    const double ud = (double)R.r * dinv;
    const bool clipped = ud >= 1e8;
    const T u = clipped ? (T)1e8 : (T)ud;
    const T au = a * u;
    T tc = 1, td0 = 1, tq0 = 1, tq1 = 1, sc = 0, sd0 = 0, sq0 = 0, sq1 = 0;
    if (au < (T)50) {                                      // pme.py:418 (expau := 0 beyond)
        const T e = exp(-au), au2 = au * au, au3 = au2 * au, au4 = au2 * au2;
        const T base = (T)1 + au + (T)0.5 * au2;
        tc = (T)1 - e * base;
        td0 = (T)1 - e * (base + (T)0.25 * au3);
        tq1 = (T)1 - e * (base + au3 * (T)(1.0 / 6));
        tq0 = (T)1 - e * (base + au3 * (T)(1.0 / 6) + au4 * (T)(1.0 / 18));
        sc = e * au2 * (T)0.5;
        sd0 = e * (au3 - au2) * (T)0.25;
        sq0 = e * (au4 - au3) * (T)(1.0 / 18);
        sq1 = e * au3 * (T)(1.0 / 6);
    }
    const T au_r = clipped ? (T)0 : a * (T)dinv;
    C.au_a = u;
    C.au_d = clipped ? (T)0 : -a * R.r * (T)(dinv * dinv);
    const T d3 = R.dxnX(3, x2, x4), d5 = R.dxnX(5, x4, x6);
    // cud/2
    const T t1 = p * tc + R.b2;
    const T B1 = ri[2] * t1;
    // dud0/2, dud1/2
    const T th0 = 3 * (p * td0 + R.b3) + x3 * X;
    const T h0 = (T)(-2.0 / 3) * ri[3] * th0;
    const T th1_ = p * tc + R.b3 - (T)(2.0 / 3) * x3 * X;
    const T h1 = ri[3] * th1_;
    // udq0/2, udq1/2
    const T tq0_ = 3 * (p * tq0 + R.b3) + (T)(4.0 / 3) * x5 * X;
    const T q0 = ri[4] * tq0_;
    const T q1 = -s3 * ri[4] * (p * tq1 + R.b3);
    // udud0, udud1 (uscales = 1, pme.py:470)
    const T tu0 = 3 * (td0 + R.b3) + x3 * X;
    const T u0 = (T)(-2.0 / 3) * ri[3] * tu0;
    const T u1 = ri[3] * (tc + R.b3 - (T)(2.0 / 3) * x3 * X);
    C.B[0] = B1; C.B[1] = h0 - h1; C.B[2] = h1; C.B[3] = q0 - c23 * q1; C.B[4] = c23 * q1; C.B[5] = u0 - u1; C.B[6] = u1;
    if (DERIV) {
        // partial d/dr at fixed au
        const T dB1 = -2 * rinv * B1 + ri[2] * R.db2;
        const T dh0 = -3 * rinv * h0 + (T)(-2.0 / 3) * ri[3] * (3 * R.db3 + d3);
        const T dh1 = -3 * rinv * h1 + ri[3] * (R.db3 - (T)(2.0 / 3) * d3);
        const T dq0 = -4 * rinv * q0 + ri[4] * (3 * R.db3 + (T)(4.0 / 3) * d5);
        const T dq1 = -4 * rinv * q1 - s3 * ri[4] * R.db3;
        const T du0 = -3 * rinv * u0 + (T)(-2.0 / 3) * ri[3] * (3 * R.db3 + d3);
        const T du1 = -3 * rinv * u1 + ri[3] * (R.db3 - (T)(2.0 / 3) * d3);
        // d/d(au)
        const T aB1 = ri[2] * p * sc;
        const T ah0 = -2 * ri[3] * p * sd0, ah1 = ri[3] * p * sc;
        const T aq0 = 3 * ri[4] * p * sq0, aq1 = -s3 * ri[4] * p * sq1;
        const T au0 = -2 * ri[3] * sd0, au1 = ri[3] * sc;
        C.aB[0] = aB1; C.aB[1] = ah0 - ah1; C.aB[2] = ah1; C.aB[3] = aq0 - c23 * aq1; C.aB[4] = c23 * aq1; C.aB[5] = au0 - au1; C.aB[6] = au1;
        C.dB[0] = dB1 + C.aB[0] * au_r; C.dB[1] = dh0 - dh1 + C.aB[1] * au_r; C.dB[2] = dh1 + C.aB[2] * au_r;
        C.dB[3] = dq0 - c23 * dq1 + C.aB[3] * au_r; C.dB[4] = c23 * dq1 + C.aB[4] * au_r;
        C.dB[5] = du0 - du1 + C.aB[5] * au_r; C.dB[6] = du1 + C.aB[6] * au_r;
        // d/dpscale
        const T pB1 = ri[2] * tc, ph0 = -2 * ri[3] * td0, ph1 = ri[3] * tc, pq0 = 3 * ri[4] * tq0, pq1 = -s3 * ri[4] * tq1;
        C.pB[0] = pB1; C.pB[1] = ph0 - ph1; C.pB[2] = ph1; C.pB[3] = pq0 - c23 * pq1; C.pB[4] = c23 * pq1; C.pB[5] = 0; C.pB[6] = 0;
    }
}

template <typename T> __device__ __forceinline__ T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// v = Theta n for the 6-component symmetric layout (xx,xy,xz,yy,yz,zz)
template <typename T> __device__ __forceinline__ void symv(const T* t, const T* n, T* v) {
    v[0] = t[0] * n[0] + t[1] * n[1] + t[2] * n[2];
    v[1] = t[1] * n[0] + t[3] * n[1] + t[4] * n[2];
    v[2] = t[2] * n[0] + t[4] * n[1] + t[5] * n[2];
}

// Packed per-atom record the staged (MODE 0) kernel gathers: position 3, Cartesian multipoles 10, induced dipole 3,
// polarizability, Thole width = 18 reals, padded to whole 16-byte chunks (double: 18 = 9 chunks; float: 20 = 5
// chunks). One record = 9 (5) LDGSTS.128 + 9 (5) LDS.128 per pair end instead of 18 + 18 eight-byte ones: the
// pair loop is bound by the load/store unit (cp.async + shared loads + 32 reductions per pair), not by DRAM.
template <typename T> struct PairRec {
    static constexpr int EPC = 16 / sizeof(T);                 // elements per 16-byte chunk
    static constexpr int STRIDE = 20;                          // elements per record: 18 + pol^(1/6) (slot 18) + pad
    static constexpr int chunks(bool pol) { return ((pol ? 18 : 13) + EPC - 1) / EPC; }   // what the flat kernel stages
    static constexpr int chunks_all() { return STRIDE / EPC; }                            // incl. slot 18 (cluster kernel)
};
template <typename T> struct alignas(16) PairChunk { T v[16 / sizeof(T)]; };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

// sum of 16 per-lane values over the warp: after the call lane L holds the total of value
// (L >> 1) & 15 (both lanes of a pair hold the same total). 16 shuffles instead of 80.
template <typename T>
__device__ __forceinline__ T warp_reduce_scatter16(T (&v)[16], int lane) {
    T w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const T keep = b4 ? v[8 + k] : v[k], give = b4 ? v[k] : v[8 + k];
        w8[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const T keep = b3 ? w8[4 + k] : w8[k], give = b3 ? w8[k] : w8[4 + k];
        w4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const T keep = b2 ? w4[2 + k] : w4[k], give = b2 ? w4[k] : w4[2 + k];
        w2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    const T keep = b1 ? w2[1] : w2[0], give = b1 ? w2[0] : w2[1];
    T r = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;      // value index 8*b4 + 4*b3 + 2*b2 + b1
}

}  // namespace admp
