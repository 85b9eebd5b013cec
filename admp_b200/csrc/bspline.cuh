// Order-6 cardinal B-spline weights and the per-atom mesh anchor shared by the spread / gather kernels
// (recip.cu: one warp / four lanes per atom; spread_brick.cu: one block per mesh brick). Replaces
// admp/recip.py:80-137 (bspline6 and its derivatives) and :296-311 (get_recip_vectors / u_reference).
#pragma once
#include "common.cuh"

namespace admp {

// w[p*6+k] = d^p/du^p M6(f+k), p = 0..NP-1, by the Cox-de Boor recursion (A23: any stable
// evaluation is acceptable; the reference's piecewise polynomials recip.py:80-137 agree to 4e-13)
template <typename T, int NP>
__device__ __forceinline__ void bspline6(T f, T* __restrict__ w) {
    T a[8][4];   // a[k+1][o-3] = M_o(f+k), o = 3..6, with zero guards at both ends
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int o = 0; o < 4; ++o) a[k][o] = (T)0;
    const T m2_0 = f, m2_1 = (T)1 - f;
    // order 3 from order 2
    a[1][0] = (T)0.5 * (f * m2_0);
    a[2][0] = (T)0.5 * ((f + 1) * m2_1 + ((T)2 - f) * m2_0);
    a[3][0] = (T)0.5 * (((T)1 - f) * m2_1);
#pragma unroll
    for (int o = 4; o <= 6; ++o) {
        const T inv = (T)1 / (T)(o - 1);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (k < o) {
                const T prev_k = (k < o - 1) ? a[k + 1][o - 4] : (T)0;
                const T prev_km1 = (k >= 1) ? a[k][o - 4] : (T)0;
                a[k + 1][o - 3] = ((f + (T)k) * prev_k + ((T)o - f - (T)k) * prev_km1) * inv;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        w[k] = a[k + 1][3];
        if (NP > 1) w[6 + k] = a[k + 1][2] - a[k][2];
        if (NP > 2) w[12 + k] = a[k + 1][1] - 2 * a[k][1] + (k >= 1 ? a[k - 1][1] : (T)0);
        if (NP > 3) w[18 + k] = a[k + 1][0] - 3 * a[k][0] + 3 * (k >= 1 ? a[k - 1][0] : (T)0) - (k >= 2 ? a[k - 2][0] : (T)0);
    }
}

// per-atom mesh anchor: fractional coordinate x_d = Nstar[d].r (double: keeps the f32 build from
// losing the sub-cell offset), m0 = ceil(x), f = m0 - x, first stencil index (m0 - 3) mod K
__device__ __forceinline__ void mesh_anchor(const BoxInfo& B, double rx, double ry, double rz, int d, double& f, int& i0) {
    const double x = B.nstar[3 * d] * rx + B.nstar[3 * d + 1] * ry + B.nstar[3 * d + 2] * rz;
    const double m0 = ceil(x);
    f = m0 - x;
    const int K = B.K[d];
    long long i = (long long)m0 - 3;
    i %= K;
    if (i < 0) i += K;
    i0 = (int)i;
}

}  // namespace admp
