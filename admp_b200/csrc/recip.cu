// Reciprocal-space kernels: order-6 B-spline multipole spread, influence-function
// convolution (+ energy, + k-space virial), and potential / field / field-gradient /
// force gather. Replaces admp/recip.py:21-431 and what jax.grad derives from it.
//
// All three are HBM-bound; algorithmic bytes (w = sizeof(real), G = K1*K2*K3):
//   spread   : w*G (mesh zero-fill) + Na*(3+n_comp)*w read + 216*Na RMW updates
//   convolve : 2 * (2w) * (G/2)  (half spectrum read + write, in place)
//   gather   : min(w*G, 216*w*Na) mesh reads + Na*(n_comp+3)*w
// The FFTs in between are cuFFT (D2Z/Z2D or R2C/C2R), see api.cu.
//
// Fractional formulation (tools/analytic_proto.py, validated vs the oracle):
//   u_d = m0_d - Nstar[d].r + 3 + s ; w_d^p[k] = d^p M6/du^p at f_d + k
//   mesh(g) += q W000 + sum_d muf_d W[e_d] + sum_de Tf_de W[e_d+e_e]
//   muf_d = -sum_c Nstar[d][c] mu_c ; Tf_de = sum_ab Nstar[d][a] Nstar[e][b] T_ab / 3
#include <cstdlib>

#include "kernels.h"
#include "influence.cuh"
#include "bspline.cuh"

namespace admp {

constexpr int SPREAD_WARPS = 4;

// One warp per atom; lanes stride the 216 stencil points (z fastest => 6 contiguous reals).
// PEER: the mesh is x-slab decomposed over several GPUs (PeerTab): the atomics of a stencil plane go to the
// buffer of the rank that owns the plane (NVLink peer atomics for the few planes that cross a slab boundary).
template <typename T, bool MULTIPOLE, bool PEER>
__global__ void __launch_bounds__(SPREAD_WARPS * 32)
spread_kernel(int n, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos, const T* __restrict__ M, int m_stride,
              const T* __restrict__ U, T* __restrict__ mesh, PeerTab peers, int zld) {
    __shared__ T sw[SPREAD_WARPS][3][18];
    __shared__ int si[SPREAD_WARPS][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * SPREAD_WARPS + warp;
    if (a >= n) return;
    const BoxInfo& B = *Bp;
    if (lane < 3) {
        double f; int i0;
        mesh_anchor(B, (double)pos[3 * a], (double)pos[3 * a + 1], (double)pos[3 * a + 2], lane, f, i0);
        bspline6<T, MULTIPOLE ? 3 : 1>((T)f, sw[warp][lane]);
        si[warp][lane] = i0;
    }
    // fractional multipole coefficients (all lanes, redundantly: ~60 flop)
    const T* m = M + (size_t)a * m_stride;
    const T q = m[0];
    T muf[3] = {0, 0, 0}, Tf[6] = {0, 0, 0, 0, 0, 0};
    if (MULTIPOLE) {
        T mu[3] = {m[1], m[2], m[3]};
        if (U != nullptr) { mu[0] += U[3 * a]; mu[1] += U[3 * a + 1]; mu[2] += U[3 * a + 2]; }
        const T Tm[9] = {m[4], m[5], m[6], m[5], m[7], m[8], m[6], m[8], m[9]};
        T N[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) N[k] = (T)B.nstar[k];
#pragma unroll
        for (int d = 0; d < 3; ++d) muf[d] = -(N[3 * d] * mu[0] + N[3 * d + 1] * mu[1] + N[3 * d + 2] * mu[2]);
        T NT[9];   // NT[d][b] = sum_a N[d][a] T[a][b]
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int b = 0; b < 3; ++b) NT[3 * d + b] = N[3 * d] * Tm[b] + N[3 * d + 1] * Tm[3 + b] + N[3 * d + 2] * Tm[6 + b];
        auto tf = [&](int d, int e) { return (NT[3 * d] * N[3 * e] + NT[3 * d + 1] * N[3 * e + 1] + NT[3 * d + 2] * N[3 * e + 2]) * (T)(1.0 / 3); };
        Tf[0] = tf(0, 0); Tf[1] = 2 * tf(0, 1); Tf[2] = 2 * tf(0, 2); Tf[3] = tf(1, 1); Tf[4] = 2 * tf(1, 2); Tf[5] = tf(2, 2);
    }
    __syncwarp();
    const T* w0 = sw[warp][0];
    const T* w1 = sw[warp][1];
    const T* w2 = sw[warp][2];
    const int K1 = B.K[0], K2 = B.K[1], K3 = B.K[2];
    const int zl = zld > 0 ? zld : K3;          // reals per mesh line (2 (K3/2 + 1) when the mesh lives in the spectrum buffer)
    const int i0 = si[warp][0], j0 = si[warp][1], k0 = si[warp][2];
    for (int pt = lane; pt < 216; pt += 32) {
        const int ia = pt / 36, ib = (pt / 6) % 6, ic = pt % 6;
        T val;
        if (MULTIPOLE) {
            const T a0 = w0[ia], a1 = w0[6 + ia], a2 = w0[12 + ia];
            const T b0 = w1[ib], b1 = w1[6 + ib], b2 = w1[12 + ib];
            const T c0 = w2[ic], c1 = w2[6 + ic], c2 = w2[12 + ic];
            const T t0 = q * a0 * b0 + muf[0] * a1 * b0 + muf[1] * a0 * b1 + Tf[0] * a2 * b0 + Tf[1] * a1 * b1 + Tf[3] * a0 * b2;
            const T t1 = muf[2] * a0 * b0 + Tf[2] * a1 * b0 + Tf[4] * a0 * b1;
            const T t2 = Tf[5] * a0 * b0;
            val = t0 * c0 + t1 * c1 + t2 * c2;
        } else {
            val = q * w0[ia] * w1[ib] * w2[ic];
        }
        int gi = i0 + ia; if (gi >= K1) gi -= K1;
        int gj = j0 + ib; if (gj >= K2) gj -= K2;
        int gk = k0 + ic; if (gk >= K3) gk -= K3;
        T* base = PEER ? reinterpret_cast<T*>(peers.base[peers.owner(gi)]) : mesh;
        atomicAdd(base + ((size_t)gi * K2 + gj) * zl + gk, val);
    }
}

// ---------------------------------------------------------------------------- convolution
template <typename T> struct cplx { T x, y; };

// per-evaluation separable tables for the Coulomb / orthorhombic fast path (see influence.cuh)
__global__ void conv_tables_kernel(const BoxInfo* __restrict__ Bp, double kappa, const double* __restrict__ bt1,
                                   const double* __restrict__ bt2, const double* __restrict__ bt3, double* __restrict__ ek,
                                   double* __restrict__ k2, int* __restrict__ ortho) {
    const BoxInfo& B = *Bp;
    const int K[3] = {B.K[0], B.K[1], B.K[2]};
    const int cnt[3] = {K[0], K[1], K[2] / 2 + 1};
    const double* bt[3] = {bt1, bt2, bt3};
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int off = 0;
    for (int d = 0; d < 3; ++d) {
        if (i < cnt[d]) {
            const double m = (d == 2) ? (double)i : (double)kint(i, K[d]);
            const double k = 6.283185307179586 * m * B.inv[4 * d];
            k2[off + i] = k * k;
            ek[off + i] = exp(-k * k / (4 * kappa * kappa)) * bt[d][i];
        }
        off += cnt[d];
    }
    if (i == 0) {
        bool o = true;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                if (a != b && B.box[3 * a + b] != 0.0) o = false;
        *ortho = o ? 1 : 0;
    }
}

// In place on the half spectrum S (K1 x K2 x (K3/2+1)):  S <- 2*scale*C_k/theta_k^2 * S and
// E += scale * sum_full C_k |S_k|^2 / theta_k^2   (recip.py:400-426, A16; R2C weights per plane).
template <typename T>
__global__ void __launch_bounds__(256)
convolve_kernel(const BoxInfo* __restrict__ Bp, T kappa, int kind, ConvTables tb, cplx<T>* __restrict__ S,
                double* __restrict__ scalars, int want_vir) {
    __shared__ double red[7 * 8];
    const BoxInfo& B = *Bp;
    const unsigned K2 = B.K[1], K3 = B.K[2], K3h = K3 / 2 + 1;
    const size_t total = (size_t)B.K[0] * K2 * K3h;
    const double scale = (kind == ADMP_CK_COULOMB) ? ADMP_DIEL : 1.0;
    const bool ortho = *tb.ortho != 0;
    const double kap = (double)kappa;
    double acc_e = 0.0, acc_t[6] = {0, 0, 0, 0, 0, 0};
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t rest = idx / K3h;
        const int i3 = (int)(idx - rest * K3h);
        const int i1 = (int)(rest / K2), i2 = (int)(rest - (size_t)i1 * K2);
        cplx<T> s = S[idx];
        const double s2 = (double)s.x * s.x + (double)s.y * s.y;
        const bool single = (i3 == 0) || (2 * i3 == (int)K3);
        if (want_vir) {
            const Influence f = influence<true>(B, tb, ortho, kap, kind, i1, i2, i3);
            acc_e += (single ? 1.0 : 2.0) * f.g * s2;
            virial_terms(B, f.kv, i1, i2, i3, single, f.dg * s2, acc_t);
            const T g = (T)(2.0 * scale * f.g);
            s.x *= g; s.y *= g;
        } else {
            const Influence f = influence<false>(B, tb, ortho, kap, kind, i1, i2, i3);
            acc_e += (single ? 1.0 : 2.0) * f.g * s2;
            const T g = (T)(2.0 * scale * f.g);
            s.x *= g; s.y *= g;
        }
        S[idx] = s;
    }
    double e1[1] = {acc_e * scale};
    block_accumulate<1>(e1, red, scalars + ADMP_S_E_RECIP);
    if (want_vir) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc_t[k] *= scale;
        block_accumulate<6>(acc_t, red, scalars + ADMP_S_TK);
    }
}

// ---------------------------------------------------------------------------- gather
// lanes per atom (LPA): the 36 (x, y) stencil columns are dealt over 4 lanes (9 columns each: least total work,
// large systems) or 16 lanes (2-3 columns each: a 4x shorter dependent chain per thread and 4x more blocks; the
// 3 072-atom boxes are latency-bound and fill only 96 blocks with 4 lanes per atom)
constexpr int GATHER_SMALL_N = 32768;               // below this many atoms: 16 lanes per atom

// MODE 0: everything (dE/dM, dE/dr, dE/dNstar).  MODE 1: field only (dE/dmu accumulated into F).
// Four lanes per atom, each lane owns 9 of the 36 (ia, ib) stencil columns. A column is 6 consecutive mesh
// points in z (one 48-byte segment); the stencil sum is evaluated separably:
//   S_p3 = sum_ic phi[ia, ib, ic] * d^p3 M6(z)[ic] ;  P[p1 p2 p3] += d^p1 M6(x)[ia] * d^p2 M6(y)[ib] * S_p3
// (54 FMA-class operations per column instead of ~50 per mesh point), then two shuffle steps reduce the
// 20 (4 in field mode) partial sums over the four lanes and lane 0 of the group does the per-atom algebra.
template <typename T, bool MULTIPOLE, int MODE, bool PEER, int LPA>
__global__ void __launch_bounds__(128)
gather_kernel(int n, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos, const T* __restrict__ M, int m_stride,
              const T* __restrict__ U, const T* __restrict__ phi, uint32_t flags, T* __restrict__ dpos, T* __restrict__ G,
              int g_stride, T* __restrict__ F, double* __restrict__ scalars, PeerTab peers, int zld) {
    constexpr int NP = (MODE == 1) ? 2 : (MULTIPOLE ? 4 : 2);
    constexpr int GATHER_LPA = LPA, GATHER_APB = 128 / LPA;
    __shared__ T sw[GATHER_APB][3][6 * NP];
    __shared__ int si[GATHER_APB][3];
    __shared__ double red[9 * 4];
    const int lane = threadIdx.x & 31;
    const int sub = lane & (GATHER_LPA - 1);
    const int slot = threadIdx.x / GATHER_LPA;
    const int a = blockIdx.x * GATHER_APB + slot;
    const BoxInfo& B = *Bp;
    const bool want_vir = (flags & ADMP_WANT_VIRIAL) != 0 && MODE == 0;
    double wacc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool valid = a < n;
    if (valid && sub < 3) {
        double f; int i0;
        mesh_anchor(B, (double)pos[3 * a], (double)pos[3 * a + 1], (double)pos[3 * a + 2], sub, f, i0);
        bspline6<T, NP>((T)f, sw[slot][sub]);
        si[slot][sub] = i0;
    }
    __syncwarp();
    // P index = derivative orders (p1 p2 p3) with p1+p2+p3 <= NP-1, in the order listed below
    constexpr int NC = (NP == 4) ? 20 : 4;
    T P[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) P[k] = (T)0;
    if (valid) {
        const T* w0 = sw[slot][0];
        const T* w1 = sw[slot][1];
        const T* w2 = sw[slot][2];
        const int K1 = B.K[0], K2 = B.K[1], K3 = B.K[2];
        const int i0 = si[slot][0], j0 = si[slot][1], k0 = si[slot][2];
        const bool wrap = k0 + 5 >= K3;
        const int zl = zld > 0 ? zld : K3;
#pragma unroll
        for (int c = 0; c < (36 + GATHER_LPA - 1) / GATHER_LPA; ++c) {
            const int col = sub + GATHER_LPA * c;
            if (36 % GATHER_LPA != 0 && col >= 36) break;
            const int ia = col / 6, ib = col - 6 * ia;
            int gi = i0 + ia; if (gi >= K1) gi -= K1;
            int gj = j0 + ib; if (gj >= K2) gj -= K2;
            const T* line = (PEER ? reinterpret_cast<const T*>(peers.base[peers.owner(gi)]) : phi) + ((size_t)gi * K2 + gj) * zl;
            T ph[6];
            if (!wrap) {
#pragma unroll
                for (int ic = 0; ic < 6; ++ic) ph[ic] = line[k0 + ic];
            } else {
#pragma unroll
                for (int ic = 0; ic < 6; ++ic) { int gk = k0 + ic; if (gk >= K3) gk -= K3; ph[ic] = line[gk]; }
            }
            T S[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                T s = ph[0] * w2[6 * q];
#pragma unroll
                for (int ic = 1; ic < 6; ++ic) s += ph[ic] * w2[6 * q + ic];
                S[q] = s;
            }
            if (NP == 4) {
                const T a0 = w0[ia], a1 = w0[6 + ia], a2 = w0[12 + ia], a3 = w0[18 + ia];
                const T b0 = w1[ib], b1 = w1[6 + ib], b2 = w1[12 + ib], b3 = w1[18 + ib];
                const T ab00 = a0 * b0, ab10 = a1 * b0, ab01 = a0 * b1, ab20 = a2 * b0, ab11 = a1 * b1, ab02 = a0 * b2;
                P[0] += ab00 * S[0];                                                              // 000
                P[1] += ab10 * S[0]; P[2] += ab01 * S[0]; P[3] += ab00 * S[1];                     // 100 010 001
                P[4] += ab20 * S[0]; P[5] += ab11 * S[0]; P[6] += ab10 * S[1];                     // 200 110 101
                P[7] += ab02 * S[0]; P[8] += ab01 * S[1]; P[9] += ab00 * S[2];                     // 020 011 002
                P[10] += a3 * b0 * S[0]; P[11] += a2 * b1 * S[0]; P[12] += ab20 * S[1];            // 300 210 201
                P[13] += a1 * b2 * S[0]; P[14] += ab11 * S[1]; P[15] += ab10 * S[2];               // 120 111 102
                P[16] += a0 * b3 * S[0]; P[17] += ab02 * S[1]; P[18] += ab01 * S[2]; P[19] += ab00 * S[3];   // 030 021 012 003
            } else {
                const T a0 = w0[ia], a1 = w0[6 + ia], b0 = w1[ib], b1 = w1[6 + ib];
                const T ab00 = a0 * b0;
                P[0] += ab00 * S[0]; P[1] += a1 * b0 * S[0]; P[2] += a0 * b1 * S[0]; P[3] += ab00 * S[1];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
#pragma unroll
        for (int o = 1; o < GATHER_LPA; o <<= 1) P[k] += __shfl_xor_sync(0xffffffffu, P[k], o);
    }
    if (valid && sub == 0) {
        T N[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) N[k] = (T)B.nstar[k];
        const T ph1[3] = {P[1], P[2], P[3]};
        if (MODE == 1) {
            // dE/dmu_c = -sum_d N[d][c] ph1[d]
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(F + (size_t)a * 3 + c, -(N[c] * ph1[0] + N[3 + c] * ph1[1] + N[6 + c] * ph1[2]));
        } else if (!MULTIPOLE) {
            const T q = M[(size_t)a * m_stride];
            if (G != nullptr) atomicAdd(G + (size_t)a * g_stride, P[0]);
            T dEdu[3] = {q * ph1[0], q * ph1[1], q * ph1[2]};
            if (dpos != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(dpos + (size_t)a * 3 + c, -(N[c] * dEdu[0] + N[3 + c] * dEdu[1] + N[6 + c] * dEdu[2]));
            }
            if (want_vir) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c) wacc[3 * d + c] = -(double)(dEdu[d] * pos[3 * a + c]);
            }
        } else {
            const T* m = M + (size_t)a * m_stride;
            const T q = m[0];
            T mu[3] = {m[1], m[2], m[3]};
            if (U != nullptr) { mu[0] += U[3 * a]; mu[1] += U[3 * a + 1]; mu[2] += U[3 * a + 2]; }
            const T Tm[9] = {m[4], m[5], m[6], m[5], m[7], m[8], m[6], m[8], m[9]};
            // symmetric second / third derivative tables
            const T ph2[9] = {P[4], P[5], P[6], P[5], P[7], P[8], P[6], P[8], P[9]};
            // ph3[d][e][f]: index by sorted multiset
            auto p3 = [&](int d, int e, int f) -> T {
                const int c0 = (d == 0) + (e == 0) + (f == 0), c1 = (d == 1) + (e == 1) + (f == 1);
                // (c0,c1,c2) -> slot
                if (c0 == 3) return P[10]; if (c0 == 2 && c1 == 1) return P[11]; if (c0 == 2) return P[12];
                if (c0 == 1 && c1 == 2) return P[13]; if (c0 == 1 && c1 == 1) return P[14]; if (c0 == 1) return P[15];
                if (c1 == 3) return P[16]; if (c1 == 2) return P[17]; if (c1 == 1) return P[18]; return P[19];
            };
            T muf[3], NT[9], Tf[9];
#pragma unroll
            for (int d = 0; d < 3; ++d) muf[d] = -(N[3 * d] * mu[0] + N[3 * d + 1] * mu[1] + N[3 * d + 2] * mu[2]);
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int b = 0; b < 3; ++b) NT[3 * d + b] = N[3 * d] * Tm[b] + N[3 * d + 1] * Tm[3 + b] + N[3 * d + 2] * Tm[6 + b];
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    Tf[3 * d + e] = (NT[3 * d] * N[3 * e] + NT[3 * d + 1] * N[3 * e + 1] + NT[3 * d + 2] * N[3 * e + 2]) * (T)(1.0 / 3);
            if (G != nullptr) {
                T* g = G + (size_t)a * g_stride;
                atomicAdd(g, P[0]);
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(g + 1 + c, -(N[c] * ph1[0] + N[3 + c] * ph1[1] + N[6 + c] * ph1[2]));
                // Gm = N^T ph2 N / 3
                T PN[9];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int b = 0; b < 3; ++b) PN[3 * d + b] = ph2[3 * d] * N[b] + ph2[3 * d + 1] * N[3 + b] + ph2[3 * d + 2] * N[6 + b];
                auto gm = [&](int p, int b) { return (N[p] * PN[b] + N[3 + p] * PN[3 + b] + N[6 + p] * PN[6 + b]) * (T)(1.0 / 3); };
                atomicAdd(g + 4, gm(0, 0)); atomicAdd(g + 5, 2 * gm(0, 1)); atomicAdd(g + 6, 2 * gm(0, 2));
                atomicAdd(g + 7, gm(1, 1)); atomicAdd(g + 8, 2 * gm(1, 2)); atomicAdd(g + 9, gm(2, 2));
            }
            if (F != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(F + (size_t)a * 3 + c, -(N[c] * ph1[0] + N[3 + c] * ph1[1] + N[6 + c] * ph1[2]));
            }
            T dEdu[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                T s = q * ph1[d];
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    s += ph2[3 * d + e] * muf[e];
#pragma unroll
                    for (int f = 0; f < 3; ++f) s += p3(d, e, f) * Tf[3 * e + f];
                }
                dEdu[d] = s;
            }
            if (dpos != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(dpos + (size_t)a * 3 + c, -(N[c] * dEdu[0] + N[3 + c] * dEdu[1] + N[6 + c] * dEdu[2]));
            }
            if (want_vir) {
                // W[d][c] = dEdu[d] (-r_c) + ph1[d] (-mu_c) + 2/3 sum_e ph2[d][e] (N T)[e][c]
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        wacc[3 * d + c] = (double)(-dEdu[d] * pos[3 * a + c] - ph1[d] * mu[c]
                            + (T)(2.0 / 3) * (ph2[3 * d] * NT[c] + ph2[3 * d + 1] * NT[3 + c] + ph2[3 * d + 2] * NT[6 + c]));
            }
        }
    }
    if (want_vir) block_accumulate<9>(wacc, red, scalars + ADMP_S_DNSTAR);
}

// ---------------------------------------------------------------------------- launchers
template <typename T>
void launch_spread(cudaStream_t st, int n, const BoxInfo* B, const void* pos, const void* M, int m_cols, int m_stride,
                   const void* U, void* mesh, const PeerTab* peers, int zld) {
    if (n <= 0) return;
    const unsigned grid = (n + SPREAD_WARPS - 1) / SPREAD_WARPS;
    const PeerTab pt = peers ? *peers : PeerTab{};
#define ADMP_S_ARGS(u) n, B, (const T*)pos, (const T*)M, m_stride, u, (T*)mesh, pt, zld
    if (peers) {
        if (m_cols >= 10) spread_kernel<T, true, true><<<grid, SPREAD_WARPS * 32, 0, st>>>(ADMP_S_ARGS((const T*)U));
        else spread_kernel<T, false, true><<<grid, SPREAD_WARPS * 32, 0, st>>>(ADMP_S_ARGS(nullptr));
    } else {
        if (m_cols >= 10) spread_kernel<T, true, false><<<grid, SPREAD_WARPS * 32, 0, st>>>(ADMP_S_ARGS((const T*)U));
        else spread_kernel<T, false, false><<<grid, SPREAD_WARPS * 32, 0, st>>>(ADMP_S_ARGS(nullptr));
    }
#undef ADMP_S_ARGS
}
template <typename T>
void launch_convolve(cudaStream_t st, const BoxInfo* B, size_t n_half, int n_sm, double kappa, int kind, const ConvTables& tb,
                     void* S, double* scalars, int want_vir) {
    size_t blocks = (n_half + 255) / 256;
    const size_t cap = (size_t)n_sm * 8;            // grid sized in multiples of the SM count (grid-stride loop)
    if (blocks > cap) blocks = cap;
    convolve_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(B, (T)kappa, kind, tb, (cplx<T>*)S, scalars, want_vir);
}
void launch_conv_tables(cudaStream_t st, const BoxInfo* B, double kappa, const double* bt1, const double* bt2, const double* bt3,
                        double* ek, double* k2, int* ortho, int maxK) {
    conv_tables_kernel<<<(maxK + 127) / 128, 128, 0, st>>>(B, kappa, bt1, bt2, bt3, ek, k2, ortho);
}
template <typename T>
void launch_gather(cudaStream_t st, int n, const BoxInfo* B, const void* pos, const void* M, int m_cols, int m_stride, const void* U,
                   const void* phi, int mode, uint32_t flags, void* dpos, void* G, int g_stride, void* F, double* scalars,
                   const PeerTab* peers, int zld, int narrow) {
    if (n <= 0) return;
    static const int force_lpa = [] { const char* e = getenv("ADMP_GATHER_LPA"); return e ? atoi(e) : 0; }();
    // narrow (admp_ctx_set_in_flight > 1): 4 lanes per atom whatever n - a quarter of the blocks, so that the gathers of several
    // evaluations in flight share the SMs with the other evaluations' passes (C2, four in flight: 474 -> 481 evals/s)
    const bool wide = force_lpa ? force_lpa == 16 : (n < GATHER_SMALL_N && !narrow);
    const int apb = 128 / (wide ? 16 : 4);
    const unsigned grid = (n + apb - 1) / apb;
    const PeerTab pt = peers ? *peers : PeerTab{};
#define ADMP_G_ARGS n, B, (const T*)pos, (const T*)M, m_stride, (const T*)U, (const T*)phi, flags, (T*)dpos, (T*)G, g_stride, (T*)F, scalars, pt, zld
#define ADMP_G_LAUNCH(PEER, LPA)                                                                         \
    do {                                                                                                 \
        if (mode == 1) gather_kernel<T, true, 1, PEER, LPA><<<grid, 128, 0, st>>>(ADMP_G_ARGS);          \
        else if (m_cols >= 10) gather_kernel<T, true, 0, PEER, LPA><<<grid, 128, 0, st>>>(ADMP_G_ARGS);  \
        else gather_kernel<T, false, 0, PEER, LPA><<<grid, 128, 0, st>>>(ADMP_G_ARGS);                   \
    } while (0)
    if (peers) {
        if (wide) ADMP_G_LAUNCH(true, 16); else ADMP_G_LAUNCH(true, 4);
    } else {
        if (wide) ADMP_G_LAUNCH(false, 16); else ADMP_G_LAUNCH(false, 4);
    }
#undef ADMP_G_LAUNCH
#undef ADMP_G_ARGS
}
#define ADMP_INST(T)                                                                                                              \
    template void launch_spread<T>(cudaStream_t, int, const BoxInfo*, const void*, const void*, int, int, const void*, void*,     \
                                   const PeerTab*, int);                                                                          \
    template void launch_convolve<T>(cudaStream_t, const BoxInfo*, size_t, int, double, int, const ConvTables&, void*, double*,   \
                                     int);                                                                                        \
    template void launch_gather<T>(cudaStream_t, int, const BoxInfo*, const void*, const void*, int, int, const void*, const void*, \
                                   int, uint32_t, void*, void*, int, void*, double*, const PeerTab*, int, int);
ADMP_INST(double)
ADMP_INST(float)
#undef ADMP_INST

}  // namespace admp
