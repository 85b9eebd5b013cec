// Influence functions C_k (admp/recip.py:434-462) with d C_k / d k^2, shared by the stand-alone
// convolution kernel (recip.cu) and the fused FFT+convolution kernel (fft.cu).
#pragma once
#include "common.cuh"

namespace admp {

__device__ __forceinline__ int kint(int i, int K) {      // recip.py:339: [0,1,...,-2,-1], even-K Nyquist negative
    return (2 * i < K) ? i : i - K;
}

// Separable fast path for Coulomb on an orthorhombic cell:
//   C_k / theta_k^2 = (2 pi / V) * prod_d [exp(-k_d^2/4kappa^2) / theta_d^2] / (k_1^2 + k_2^2 + k_3^2)
// ek[d][i] = exp(-k_d(i)^2 / 4 kappa^2) / theta_d(i)^2 ,  k2[d][i] = k_d(i)^2 ; built per evaluation
// (the box may change between evaluations) by conv_tables_kernel.
struct ConvTables {
    const double* ek[3];
    const double* k2[3];
    const double* bt[3];   // 1/theta_d^2 only (generic path)
    const int* ortho;      // device flag: 1 when the cell is orthorhombic
};

struct Influence {
    double g;      // C_k / theta_k^2           (what multiplies |S|^2 in the energy)
    double dg;     // dC_k/dk^2 / theta_k^2     (virial)
    double kv[3];  // k vector (only filled when WANT_K)
};

template <bool WANT_K>
__device__ __forceinline__ Influence influence(const BoxInfo& B, const ConvTables& tb, bool ortho, double kap, int kind,
                                               int i1, int i2, int i3) {
    Influence r;
    const double twopi = 6.283185307179586;
    const double V = B.vol;
    const int K1 = B.K[0], K2 = B.K[1];
    if (kind == ADMP_CK_COULOMB && ortho) {
        const double ksq = tb.k2[0][i1] + tb.k2[1][i2] + tb.k2[2][i3];
        if (i1 == 0 && i2 == 0 && i3 == 0) { r.g = 0.0; r.dg = 0.0; }      // gamma point dropped (recip.py:416)
        else {
            const double inv = 1.0 / ksq;
            r.g = twopi / V * tb.ek[0][i1] * tb.ek[1][i2] * tb.ek[2][i3] * inv;
            r.dg = -r.g * (inv + 1.0 / (4 * kap * kap));
        }
        if (WANT_K) {
            r.kv[0] = twopi * kint(i1, K1) * B.inv[0]; r.kv[1] = twopi * kint(i2, K2) * B.inv[4]; r.kv[2] = twopi * i3 * B.inv[8];
        }
        return r;
    }
    const double m1 = kint(i1, K1), m2 = kint(i2, K2), m3 = i3;
    double kv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) kv[c] = twopi * (m1 * B.inv[c] + m2 * B.inv[3 + c] + m3 * B.inv[6 + c]);   // recip.py:360
    const double ksq = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
    double C, dC;
    if (kind == ADMP_CK_COULOMB) {                                  // Ck_1, recip.py:434
        if (i1 == 0 && i2 == 0 && i3 == 0) { C = 0.0; dC = 0.0; }
        else {
            C = twopi / (V * ksq) * exp(-ksq / (4 * kap * kap));
            dC = -C * (1.0 / ksq + 1.0 / (4 * kap * kap));
        }
    } else {                                                        // Ck_6/8/10, recip.py:437-462
        const double x2 = ksq / (4 * kap * kap), x = sqrt(x2), e = exp(-x2), ec = ADMP_SQRT_PI * erfc(x);
        double f, df, pref;
        const double base = ADMP_SQRT_PI * 3.141592653589793 / 2 / V;
        if (kind == ADMP_CK_DISP6) {
            f = (1 - 2 * x2) * e + 2 * x2 * x * ec; df = -6 * e + 6 * x * ec; pref = base * kap * kap * kap / 3;
        } else if (kind == ADMP_CK_DISP8) {
            f = (3 - 2 * x2 + 4 * x2 * x2) * e - 4 * x2 * x2 * x * ec; df = e * (-10 + 20 * x2) - 20 * x2 * x * ec;
            pref = base * kap * kap * kap * kap * kap / 45;
        } else {
            f = (15 - 6 * x2 + 4 * x2 * x2 - 8 * x2 * x2 * x2) * e + 8 * x2 * x2 * x2 * x * ec;
            df = e * (-42 + 28 * x2 - 56 * x2 * x2) + 56 * x2 * x2 * x * ec;
            pref = base * kap * kap * kap * kap * kap * kap * kap / 1260;
        }
        C = pref * f; dC = pref * df / (8 * kap * kap);
    }
    const double th = tb.bt[0][i1] * tb.bt[1][i2] * tb.bt[2][i3];   // 1/theta_k^2, recip.py:400-408
    r.g = C * th; r.dg = dC * th;
    if (WANT_K) { r.kv[0] = kv[0]; r.kv[1] = kv[1]; r.kv[2] = kv[2]; }
    return r;
}

// k_a k_c summed over a half-spectrum point and (for weight-2 points) its Hermitian partner, whose
// k is -k except in a dimension where the point sits on an even-K Nyquist index (aliases onto itself).
__device__ __forceinline__ void virial_terms(const BoxInfo& B, const double (&kv)[3], int i1, int i2, int i3, bool single, double b,
                                             double (&acc)[6]) {
    const double twopi = 6.283185307179586;
    const int K1 = B.K[0], K2 = B.K[1];
    const double p0 = (2 * i1 == K1) ? 1.0 : -1.0, p1 = (2 * i2 == K2) ? 1.0 : -1.0;
    const double m1 = kint(i1, K1), m2 = kint(i2, K2), m3 = i3;
    double kq[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) kq[c] = twopi * (p0 * m1 * B.inv[c] + p1 * m2 * B.inv[3 + c] - m3 * B.inv[6 + c]);
    const double pw = single ? 0.0 : 1.0;
    acc[0] += b * (kv[0] * kv[0] + pw * kq[0] * kq[0]);
    acc[1] += b * (kv[0] * kv[1] + pw * kq[0] * kq[1]);
    acc[2] += b * (kv[0] * kv[2] + pw * kq[0] * kq[2]);
    acc[3] += b * (kv[1] * kv[1] + pw * kq[1] * kq[1]);
    acc[4] += b * (kv[1] * kv[2] + pw * kq[1] * kq[2]);
    acc[5] += b * (kv[2] * kv[2] + pw * kq[2] * kq[2]);
}

}  // namespace admp
