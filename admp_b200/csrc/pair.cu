// Real-space pair kernels.
//
// pme_pair_kernel replaces admp/pme.py:258-729 (calc_e_perm, calc_e_ind, pme_real_kernel,
// pme_real) AND everything jax.grad derives from them: one pass over the pair list gives
// the energy, dE/dr, the image-shift box term, dE/dM (potential / field / field gradient
// per site), dE/dU, and the scale / Thole / polarizability gradients.
//
// Formulation (validated against the oracle's autograd in tests/test_analytic_proto.py):
// instead of rotating both sites into the reference's quasi-internal frame (two 5x5
// D-matrix builds per pair), the pair energy is written with rotational invariants of
// (n = dr/|dr|, mu, Theta, u):
//   E = A0 qI qJ + A1 (qI dJ - dI qJ) + A2 dI dJ + A3 muI.muJ + A4 (tI qJ + qI tJ)
//     + A5 (tI dJ - dI tJ) + A6 (muJ.vI - muI.vJ) + A7 tI tJ + A8 vI.vJ + A9 TI:TJ
//     + B1 (qI pJ - pI qJ) + B2 (pI dJ + pJ dI) + B3 (uI.muJ + uJ.muI)
//     + B5 (tI pJ - pI tJ) + B6 (uJ.vI - uI.vJ) + C2 pI pJ + C3 uI.uJ
// with d = mu.n, p = u.n, v = Theta n, t = n.v and radial functions that are linear
// combinations of the reference's cc..qq_m2 / cud..udud_m1. FP64/FP32 CUDA-core work
// (erf, exp, rsqrt): no tensor cores, this is not a dense contraction.
#include "kernels.h"
#include "pair_math.cuh"

namespace admp {


// ------------------------------------------------------------------------------------------
// Scale index of every pair row (admp/pme.py:681-683), -1 for rows that are not evaluated (padding, i >= j):
// computed once per pair list so the 30+ pair-kernel launches of an SCF evaluation do not walk the covalent
// CSR again.
__global__ void __launch_bounds__(256)
pair_scale_kernel(int64_t n_rows, int n_atoms, const int32_t* __restrict__ pairs, const int32_t* __restrict__ cov_off,
                  const int32_t* __restrict__ cov_idx, const int8_t* __restrict__ cov_nb, int8_t* __restrict__ sidx) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rows) return;
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    const bool live = (i < j) && (i >= 0) && (j < n_atoms);              // pme.py:671 (padding rows are (N,N))
    sidx[p] = live ? (int8_t)scale_index(cov_off, cov_idx, cov_nb, i, j) : (int8_t)-1;
}
void launch_pair_scale(cudaStream_t st, int64_t n_rows, int n_atoms, const int32_t* pairs, const int32_t* cov_off,
                       const int32_t* cov_idx, const int8_t* cov_nb, int8_t* sidx) {
    if (n_rows <= 0) return;
    pair_scale_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(n_rows, n_atoms, pairs, cov_off, cov_idx, cov_nb, sidx);
}

// ------------------------------------------------------------------------------------------
// pme_pair_kernel: one pair per thread, persistent blocks, tiles of 128 pair rows.
//
//  * Neighbour tiles are staged in shared memory: the per-atom records of both ends of a pair (position,
//    Cartesian multipoles, induced dipole, polarizability, Thole width: 18 reals per atom) are gathered with
//    cp.async one tile AHEAD of the arithmetic (field-major layout [field][thread], conflict free), and the pair
//    indices two tiles ahead, so the ~1 200 FP64 operations of a pair never wait on a dependent gather.
//  * Pair lists arrive grouped by their larger index (admp_nblist_build order, like jax_md's): when all 32
//    rows of a warp share the same j, the j-side gradient (dE/dr_j, dE/dM_j, dE/dU_j: 16 values) is
//    summed over the warp with a 16-shuffle reduce-scatter and written with one atomic per value instead
//    of 16 atomics per lane; any other ordering falls back to per-lane atomics (correct for every list).
//  * MODE 0: energy + all adjoints. MODE 1: dE/dU only (the SCF field).
constexpr int PAIR_TILE = 128;
#ifndef ADMP_PAIR_MINBLOCKS
#define ADMP_PAIR_MINBLOCKS 2     // resident blocks per SM of the float64 energy + adjoint kernel (252 registers)
#endif
template <typename T>
__global__ void __launch_bounds__(256)
pair_pack_kernel(int n, const T* __restrict__ pos, const T* __restrict__ M, const T* __restrict__ U, const T* __restrict__ pol,
                 const T* __restrict__ tholes, T* __restrict__ rec) {
    constexpr int S = PairRec<T>::STRIDE;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * S) return;
    const int a = (int)(e / S), f = (int)(e - (int64_t)a * S);
    T v = (T)0;
    if (f < 3) v = pos[3 * (size_t)a + f];
    else if (f < 13) v = M[10 * (size_t)a + (f - 3)];
    else if (f < 16) v = U ? U[3 * (size_t)a + (f - 13)] : (T)0;
    else if (f == 16) v = pol ? pol[a] : (T)0;
    else if (f == 17) v = tholes ? tholes[a] : (T)0;
    else if (f == 18) v = (pol && pol[a] > (T)0) ? (T)pow((double)pol[a], 1.0 / 6.0) : (T)0;
    rec[e] = v;
}
template <typename T>
void launch_pair_pack(cudaStream_t st, int n, const void* pos, const void* M, const void* U, const void* pol, const void* tholes, void* rec) {
    if (n <= 0) return;
    const int64_t tot = (int64_t)n * PairRec<T>::STRIDE;
    pair_pack_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, (const T*)pos, (const T*)M, (const T*)U, (const T*)pol,
                                                                        (const T*)tholes, (T*)rec);
}
template void launch_pair_pack<double>(cudaStream_t, int, const void*, const void*, const void*, const void*, const void*, void*);
template void launch_pair_pack<float>(cudaStream_t, int, const void*, const void*, const void*, const void*, const void*, void*);
size_t pair_record_bytes(int dtype_bytes) { return dtype_bytes == 8 ? PairRec<double>::STRIDE * 8 : PairRec<float>::STRIDE * 4; }


template <typename T, bool POL, int MODE>
__global__ void __launch_bounds__(PAIR_TILE, (MODE == 1 || sizeof(T) == 4) ? 4 : ADMP_PAIR_MINBLOCKS)
pme_pair_kernel(int64_t n_rows, int n_atoms, const BoxInfo* __restrict__ Bp, T kappa,
                const T* __restrict__ pos, const int32_t* __restrict__ pairs, const int8_t* __restrict__ sidx_rows,
                const int32_t* __restrict__ cov_off, const int32_t* __restrict__ cov_idx, const int8_t* __restrict__ cov_nb,
                const T* __restrict__ M, const T* __restrict__ U, const T* __restrict__ pol, const T* __restrict__ tholes,
                const T* __restrict__ mScales, const T* __restrict__ pScales, uint32_t flags,
                T* __restrict__ dpos, T* __restrict__ G, T* __restrict__ F, T* __restrict__ dpol, T* __restrict__ dth,
                double* __restrict__ scalars, const T* __restrict__ rec, const int32_t* __restrict__ sel) {
    if (sel != nullptr && sel[2] != 0) return;                  // the cluster kernel (pair_cluster.cu) owns this list
    constexpr int NCH = PairRec<T>::chunks(POL), EPC = PairRec<T>::EPC, REC = PairRec<T>::STRIDE;
    // the field-only SCF kernel does ~200 FP64 operations per pair: staging 36 values per pair through shared
    // memory costs more than it hides there, so MODE 1 gathers straight from global memory (L2-resident arrays)
    constexpr bool STAGED = (MODE == 0);
    extern __shared__ __align__(16) unsigned char pair_smem[];
    __shared__ double red[10 * 4];
    __shared__ BoxInfo sB;                                      // cell + scale tables: shared-memory reads in the pair loop
    __shared__ T sScale[10], sW0[5];
    PairChunk<T>* stage_base = reinterpret_cast<PairChunk<T>*>(pair_smem);   // [2 stages][2 ends][NCH][PAIR_TILE] 16-byte chunks
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < (int)(sizeof(BoxInfo) / sizeof(double))) reinterpret_cast<double*>(&sB)[tid] = reinterpret_cast<const double*>(Bp)[tid];
    if (tid < 5) { sScale[tid] = mScales[tid]; sScale[5 + tid] = POL ? pScales[tid] : (T)0; sW0[tid] = POL ? thole_switch_w0<T>(pScales[tid]) : (T)0; }
    __syncthreads();
    double acc_e = 0.0;
    double acc_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double acc_ms[5] = {0, 0, 0, 0, 0}, acc_ps[5] = {0, 0, 0, 0, 0};
    const bool want_grad = (flags & ADMP_WANT_GRAD) != 0, want_vir = (flags & ADMP_WANT_VIRIAL) != 0,
               want_pg = (flags & ADMP_WANT_PGRAD) != 0;
    // measurement switch (tools/pair_roofline.py): drop the per-pair i-side reductions to see what they cost
    const bool iside = (flags & 0x40000000u) == 0;
    const BoxInfo& B = sB;
    const int64_t ntiles = (n_rows + PAIR_TILE - 1) / PAIR_TILE;

    struct Row { int i, j, s; };                   // s < 0: row not evaluated
    // volatile asm loads: the compiler must issue them here, two tiles ahead of their use, instead of sinking
    // them next to the consumer to shorten live ranges (which would put the load latency back on the critical path)
    auto load_row = [&](int64_t tile) -> Row {
        Row r = {0, 0, -1};
        const int64_t p = tile * PAIR_TILE + tid;
        if (tile < ntiles && p < n_rows) {
            asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];" : "=r"(r.i), "=r"(r.j) : "l"(pairs + 2 * p));
            if (sidx_rows != nullptr) asm volatile("ld.global.nc.s8 %0, [%1];" : "=r"(r.s) : "l"(sidx_rows + p));
            else if ((r.i < r.j) && (r.i >= 0) && (r.j < n_atoms)) r.s = scale_index(cov_off, cov_idx, cov_nb, r.i, r.j);
        }
        return r;
    };
    auto slot = [&](int stage, int end, int ch) -> PairChunk<T>* {
        return stage_base + ((size_t)((stage * 2 + end) * NCH + ch)) * PAIR_TILE + tid;
    };
    auto issue = [&](const Row& r, int stage) {
        if (!STAGED) return;
        if (r.s >= 0) {
#pragma unroll
            for (int end = 0; end < 2; ++end) {
                const PairChunk<T>* src = reinterpret_cast<const PairChunk<T>*>(rec + (size_t)(end == 0 ? r.i : r.j) * REC);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) cp_async16(slot(stage, end, ch), src + ch);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    int64_t tile = blockIdx.x;
    Row cur = load_row(tile);
    Row nxt = load_row(tile + gridDim.x);
    issue(cur, 0);
    int stage = 0;
    for (; tile < ntiles; tile += gridDim.x, stage ^= 1) {
        issue(nxt, stage ^ 1);                                        // gathers of the next tile
        const Row after = load_row(tile + 2 * (int64_t)gridDim.x);   // pair indices two tiles ahead
        if (STAGED) asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // this tile's records have landed (thread-private slots)
        const bool live = cur.s >= 0;
        const int i = cur.i, j = cur.j, sidx = live ? cur.s : 4;
        // j-side gradient of this pair: dE/dr_j (3), dE/dM_j (10), dE/dU_j (3)
        T gj[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) gj[k] = (T)0;
        // field f of pair end `end`: 0-2 position, 3-12 multipoles, 13-15 induced dipole, 16 polarizability, 17 Thole width
        T ra[2][STAGED ? NCH * EPC : 1];          // staged records of both ends (registers after unrolling)
        if (STAGED && live) {
#pragma unroll
            for (int end = 0; end < 2; ++end)
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const PairChunk<T> c = *slot(stage, end, ch);
#pragma unroll
                    for (int e = 0; e < EPC; ++e) ra[end][ch * EPC + e] = c.v[e];
                }
        }
        auto rd = [&](int end, int f) -> T {
            if (STAGED) return ra[end][f];
            const size_t a = end == 0 ? i : j;
            if (f < 3) return pos[3 * a + f];
            if (f < 13) return M[10 * a + (f - 3)];
            if (f < 16) return U[3 * a + (f - 13)];
            return f == 16 ? pol[a] : tholes[a];
        };
        if (live) {
            T ri_[3], rj_[3], mi[10], mj[10];
#pragma unroll
            for (int k = 0; k < 3; ++k) { ri_[k] = rd(0, k); rj_[k] = rd(1, k); }
#pragma unroll
            for (int k = 0; k < 10; ++k) { mi[k] = rd(0, 3 + k); mj[k] = rd(1, 3 + k); }
            T d[3] = {ri_[0] - rj_[0], ri_[1] - rj_[1], ri_[2] - rj_[2]};
            T sh[3];
            min_image(B, d, sh);
            const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            const T rinv = rsqrt(r2), r = r2 * rinv;
            const T n[3] = {d[0] * rinv, d[1] * rinv, d[2] * rinv};
            Radial<T> R;
            radial_setup(r, rinv, kappa, R);
            const T qI = mi[0], qJ = mj[0];
            const T *muI = mi + 1, *muJ = mj + 1, *TI = mi + 4, *TJ = mj + 4;
            T vI[3], vJ[3];
            symv(TI, n, vI); symv(TJ, n, vJ);
            const T dI = dot3(muI, n), dJ = dot3(muJ, n), tI = dot3(vI, n), tJ = dot3(vJ, n);
            T uI[3] = {0, 0, 0}, uJ[3] = {0, 0, 0}, pI = 0, pJ = 0;
            T polI = 0, polJ = 0;
            IndCoef<T> C;
            if (POL) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { uI[k] = rd(0, 13 + k); uJ[k] = rd(1, 13 + k); }
                pI = dot3(uI, n); pJ = dot3(uJ, n);
                polI = rd(0, 16); polJ = rd(1, 16);
                ind_coeffs<T, MODE == 0>(R, sScale[5 + sidx], sW0[sidx], rd(0, 17), rd(1, 17), polI, polJ, C);
            }
            if (MODE == 1) {
                // dE/du only
                const T B1 = C.B[0], B2 = C.B[1], B3 = C.B[2], B5 = C.B[3], B6 = C.B[4], C2 = C.B[5], C3 = C.B[6];
                const T e_pI = -B1 * qJ + B2 * dJ - B5 * tJ + C2 * pJ;
                const T e_pJ = B1 * qI + B2 * dI + B5 * tI + C2 * pI;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (iside) atomicAdd(F + (size_t)i * 3 + k, e_pI * n[k] + B3 * muJ[k] - B6 * vJ[k] + C3 * uJ[k]);
                    gj[13 + k] = e_pJ * n[k] + B3 * muI[k] + B6 * vI[k] + C3 * uI[k];
                }
            } else {
                T A[10], dA[10], mA[10];
                perm_coeffs<T, true>(R, sScale[sidx], A, dA, mA);
                const T mm = dot3(muI, muJ), gJI = dot3(muJ, vI), gIJ = dot3(muI, vJ), vv = dot3(vI, vJ);
                const T TT = TI[0] * TJ[0] + TI[3] * TJ[3] + TI[5] * TJ[5] + 2 * (TI[1] * TJ[1] + TI[2] * TJ[2] + TI[4] * TJ[4]);
                const T inv[10] = {qI * qJ, qI * dJ - dI * qJ, dI * dJ, mm, tI * qJ + qI * tJ, tI * dJ - dI * tJ, gJI - gIJ, tI * tJ, vv, TT};
                T e = 0, dEdr = 0, dm = 0;
#pragma unroll
                for (int k = 0; k < 10; ++k) { e += A[k] * inv[k]; dEdr += dA[k] * inv[k]; dm += mA[k] * inv[k]; }
                T e_dI = -A[1] * qJ + A[2] * dJ - A[5] * tJ, e_dJ = A[1] * qI + A[2] * dI + A[5] * tI;
                T e_tI = A[4] * qJ + A[5] * dJ + A[7] * tJ, e_tJ = A[4] * qI - A[5] * dI + A[7] * tI;
                T g_qI = A[0] * qJ + A[1] * dJ + A[4] * tJ, g_qJ = A[0] * qI - A[1] * dI + A[4] * tI;
                T e_pI = 0, e_pJ = 0;
                T inv2[7];
                if (POL) {
                    inv2[0] = qI * pJ - pI * qJ; inv2[1] = pI * dJ + pJ * dI; inv2[2] = dot3(uI, muJ) + dot3(uJ, muI);
                    inv2[3] = tI * pJ - pI * tJ; inv2[4] = dot3(uJ, vI) - dot3(uI, vJ); inv2[5] = pI * pJ; inv2[6] = dot3(uI, uJ);
#pragma unroll
                    for (int k = 0; k < 7; ++k) { e += C.B[k] * inv2[k]; dEdr += C.dB[k] * inv2[k]; }
                    const T B1 = C.B[0], B2 = C.B[1], B5 = C.B[3], C2 = C.B[5];
                    e_dI += B2 * pJ; e_dJ += B2 * pI; e_tI += B5 * pJ; e_tJ -= B5 * pI; g_qI += B1 * pJ; g_qJ -= B1 * pI;
                    e_pI = -B1 * qJ + B2 * dJ - B5 * tJ + C2 * pJ;
                    e_pJ = B1 * qI + B2 * dI + B5 * tI + C2 * pI;
                }
                acc_e += (double)e;
                if (want_pg) {
                    acc_ms[sidx] += (double)dm;
                    if (POL) {
                        T dp = 0, da = 0;
#pragma unroll
                        for (int k = 0; k < 7; ++k) { dp += C.pB[k] * inv2[k]; da += C.aB[k] * inv2[k]; }
                        acc_ps[sidx] += (double)dp;
                        const T e_th = da * C.au_a * C.da_dth;
                        if (dth != nullptr) { atomicAdd(dth + i, e_th); atomicAdd(dth + j, e_th); }
                        if (!C.trimmed && dpol != nullptr) {
                            const T e_dmp = da * C.au_d * C.dmp * (T)(1.0 / 6);
                            atomicAdd(dpol + i, e_dmp / polI); atomicAdd(dpol + j, e_dmp / polJ);
                        }
                    }
                }
                if (want_grad) {
                    const T B3 = POL ? C.B[2] : (T)0, B6 = POL ? C.B[4] : (T)0, C3 = POL ? C.B[6] : (T)0;
                    T TImuJ[3], TJmuI[3], TIvJ[3], TJvI[3];
                    symv(TI, muJ, TImuJ); symv(TJ, muI, TJmuI); symv(TI, vJ, TIvJ); symv(TJ, vI, TJvI);
                    T gn[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        gn[k] = e_dI * muI[k] + e_dJ * muJ[k] + 2 * e_tI * vI[k] + 2 * e_tJ * vJ[k] + A[6] * (TImuJ[k] - TJmuI[k]) + A[8] * (TIvJ[k] + TJvI[k]);
                    if (POL) {
                        T TIuJ[3], TJuI[3];
                        symv(TI, uJ, TIuJ); symv(TJ, uI, TJuI);
#pragma unroll
                        for (int k = 0; k < 3; ++k) gn[k] += e_pI * uI[k] + e_pJ * uJ[k] + B6 * (TIuJ[k] - TJuI[k]);
                    }
                    const T gnn = dot3(gn, n);
                    T fv[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        fv[k] = dEdr * n[k] + (gn[k] - gnn * n[k]) * rinv;
                        if (iside) atomicAdd(dpos + (size_t)i * 3 + k, fv[k]);
                        gj[k] = -fv[k];
                    }
                    if (want_vir) {
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int b = 0; b < 3; ++b) acc_box[3 * a + b] -= (double)(sh[a] * fv[b]);
                    }
                    // dE/dM
                    T* Gi = G + (size_t)i * 10;
                    if (iside) atomicAdd(Gi, g_qI);
                    gj[3] = g_qJ;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (iside) atomicAdd(Gi + 1 + k, e_dI * n[k] + A[3] * muJ[k] - A[6] * vJ[k] + B3 * uJ[k]);
                        gj[4 + k] = e_dJ * n[k] + A[3] * muI[k] + A[6] * vI[k] + B3 * uI[k];
                    }
                    // quadrupole gradient: e_t n n^T + w n^T (symmetrised into the 6-comp layout) + A9 T_other
                    T wI[3], wJ[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        wI[k] = A[6] * muJ[k] + A[8] * vJ[k] + B6 * uJ[k];
                        wJ[k] = -A[6] * muI[k] + A[8] * vI[k] - B6 * uI[k];
                    }
                    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int a = ia[k], b = ib[k];
                        const T mult = (a == b) ? (T)1 : (T)2;
                        T gI = e_tI * n[a] * n[b] * mult + ((a == b) ? wI[a] * n[a] : wI[a] * n[b] + wI[b] * n[a]) + A[9] * TJ[k] * mult;
                        T gJ = e_tJ * n[a] * n[b] * mult + ((a == b) ? wJ[a] * n[a] : wJ[a] * n[b] + wJ[b] * n[a]) + A[9] * TI[k] * mult;
                        if (iside) atomicAdd(Gi + 4 + k, gI);
                        gj[7 + k] = gJ;
                    }
                    if (POL) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (iside) atomicAdd(F + (size_t)i * 3 + k, e_pI * n[k] + B3 * muJ[k] - B6 * vJ[k] + C3 * uJ[k]);
                            gj[13 + k] = e_pJ * n[k] + B3 * muI[k] + B6 * vI[k] + C3 * uI[k];
                        }
                    }
                }
            }
        }
        // j side: one warp-wide sum when the whole warp shares j, per-lane atomics otherwise
        if (MODE == 1 || want_grad) {
            const int j0 = __shfl_sync(0xffffffffu, j, 0);
            const bool uniform = __all_sync(0xffffffffu, live && j == j0);
            if (uniform) {
                if (MODE == 1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const T s = warp_sum(gj[13 + k]);
                        if (lane == k) atomicAdd(F + (size_t)j0 * 3 + k, s);
                    }
                } else {
                    const T s = warp_reduce_scatter16(gj, lane);
                    const int v = (lane >> 1) & 15;
                    if ((lane & 1) == 0) {
                        if (v < 3) atomicAdd(dpos + (size_t)j0 * 3 + v, s);
                        else if (v < 13) atomicAdd(G + (size_t)j0 * 10 + (v - 3), s);
                        else if (POL) atomicAdd(F + (size_t)j0 * 3 + (v - 13), s);
                    }
                }
            } else if (live) {
                if (MODE == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(dpos + (size_t)j * 3 + k, gj[k]);
#pragma unroll
                    for (int k = 0; k < 10; ++k) atomicAdd(G + (size_t)j * 10 + k, gj[3 + k]);
                }
                if (POL) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(F + (size_t)j * 3 + k, gj[13 + k]);
                }
            }
        }
        cur = nxt;
        nxt = after;
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (MODE == 0) {
        double v1[1] = {acc_e};
        block_accumulate<1>(v1, red, scalars + ADMP_S_E_REAL);
        if (want_vir && want_grad) block_accumulate<9>(acc_box, red, scalars + ADMP_S_DBOX);
        if (want_pg) {
            block_accumulate<5>(acc_ms, red, scalars + ADMP_S_DMSCALE);
            if (POL) block_accumulate<5>(acc_ps, red, scalars + ADMP_S_DPSCALE);
        }
    }
}

template <typename T, bool POL, int MODE>
static void launch_pme_pair_t(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, double kappa, const void* pos,
                              const int32_t* pairs, const int8_t* sidx, const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb,
                              const void* M, const void* U, const void* pol, const void* tholes, const void* mS, const void* pS,
                              uint32_t flags, void* dpos, void* G, void* F, void* dpol, void* dth, double* scalars, const void* rec,
                              const int32_t* sel) {
    static int grid_cap = 0;
    const size_t smem = MODE == 0 ? (size_t)2 * 2 * PairRec<T>::chunks(POL) * PAIR_TILE * 16 : 0;
    auto kern = pme_pair_kernel<T, POL, MODE>;
    if (grid_cap == 0) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 1, dev = 0, nsm = 148;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PAIR_TILE, smem);
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        grid_cap = nsm * (occ > 0 ? occ : 1);        // persistent grid: a multiple of the SM count
    }
    const int64_t ntiles = (n_rows + PAIR_TILE - 1) / PAIR_TILE;
    const unsigned grid = (unsigned)(ntiles < grid_cap ? ntiles : grid_cap);
    kern<<<grid, PAIR_TILE, smem, st>>>(n_rows, n_atoms, B, (T)kappa, (const T*)pos, pairs, sidx, cov_off, cov_idx, cov_nb, (const T*)M,
                                        (const T*)U, (const T*)pol, (const T*)tholes, (const T*)mS, (const T*)pS, flags, (T*)dpos, (T*)G,
                                        (T*)F, (T*)dpol, (T*)dth, scalars, (const T*)rec, sel);
}

template <typename T>
void launch_pme_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, double kappa, const void* pos,
                     const int32_t* pairs, const int8_t* sidx, const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb,
                     const void* M, const void* U, const void* pol, const void* tholes, const void* mS, const void* pS,
                     int mode, uint32_t flags, void* dpos, void* G, void* F, void* dpol, void* dth, double* scalars, void* rec,
                     const int32_t* sel) {
    if (n_rows <= 0) return;
    const bool polz = (U != nullptr);
    // mode 0 gathers packed per-atom records (rec: n_atoms * pair_record_bytes workspace, filled here)
    if (mode != 1) launch_pair_pack<T>(st, n_atoms, pos, M, U, pol, tholes, rec);
#define ADMP_PAIR_ARGS st, n_rows, n_atoms, B, kappa, pos, pairs, sidx, cov_off, cov_idx, cov_nb, M, U, pol, tholes, mS, pS, flags, dpos, G, F, \
    dpol, dth, scalars, rec, sel
    if (mode == 1) {
        if (polz) launch_pme_pair_t<T, true, 1>(ADMP_PAIR_ARGS);
    } else if (polz) {
        launch_pme_pair_t<T, true, 0>(ADMP_PAIR_ARGS);
    } else {
        launch_pme_pair_t<T, false, 0>(ADMP_PAIR_ARGS);
    }
#undef ADMP_PAIR_ARGS
}
template void launch_pme_pair<double>(cudaStream_t, int64_t, int, const BoxInfo*, double, const void*, const int32_t*, const int8_t*,
                                      const int32_t*, const int32_t*, const int8_t*, const void*, const void*, const void*, const void*,
                                      const void*, const void*, int, uint32_t, void*, void*, void*, void*, void*, double*, void*, const int32_t*);
template void launch_pme_pair<float>(cudaStream_t, int64_t, int, const BoxInfo*, double, const void*, const int32_t*, const int8_t*,
                                     const int32_t*, const int32_t*, const int8_t*, const void*, const void*, const void*, const void*,
                                     const void*, const void*, int, uint32_t, void*, void*, void*, void*, void*, double*, void*, const int32_t*);

// ------------------------------------------------------------------------------------------
// Dispersion real space: admp/disp_pme.py:126-251.  E = sum_p (m + g_p(x^2) - 1) ci cj / r^p.
template <typename T>
__global__ void __launch_bounds__(128)
disp_pair_kernel(int64_t n_rows, int n_atoms, const BoxInfo* __restrict__ Bp, T kappa, int pmax,
                 const T* __restrict__ pos, const int32_t* __restrict__ pairs,
                 const int32_t* __restrict__ cov_off, const int32_t* __restrict__ cov_idx, const int8_t* __restrict__ cov_nb,
                 const T* __restrict__ c_list, const T* __restrict__ mScales, uint32_t flags,
                 T* __restrict__ dpos, T* __restrict__ dc, double* __restrict__ scalars) {
    __shared__ double red[10 * 4];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc_e = 0.0, acc_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, acc_ms[5] = {0, 0, 0, 0, 0};
    const bool want_grad = (flags & ADMP_WANT_GRAD) != 0, want_vir = (flags & ADMP_WANT_VIRIAL) != 0,
               want_pg = (flags & ADMP_WANT_PGRAD) != 0;
    int i = 0, j = 0;
    bool live = false;
    if (p < n_rows) { i = pairs[2 * p]; j = pairs[2 * p + 1]; live = (i < j) && (i >= 0) && (j < n_atoms); }
    if (live) {
        T d[3] = {pos[3 * i] - pos[3 * j], pos[3 * i + 1] - pos[3 * j + 1], pos[3 * i + 2] - pos[3 * j + 2]};
        T sh[3];
        min_image(*Bp, d, sh);
        const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const int sidx = scale_index(cov_off, cov_idx, cov_nb, i, j);
        const T m = mScales[sidx];
        const T x2 = kappa * kappa * r2, x4 = x2 * x2, ex = exp(-x2);
        const T ir2 = (T)1 / r2, ir6 = ir2 * ir2 * ir2;
        // g_p and dg_p/dx2 = -exp(-x2) x2^(p/2-1)/(p/2-1)!
        const T g6 = ((T)1 + x2 + (T)0.5 * x4) * ex, g8 = g6 + x4 * x2 * (T)(1.0 / 6) * ex, g10 = g8 + x4 * x4 * (T)(1.0 / 24) * ex;
        const T h6 = -ex * x4 * (T)0.5, h8 = -ex * x4 * x2 * (T)(1.0 / 6), h10 = -ex * x4 * x4 * (T)(1.0 / 24);
        const T ci[3] = {c_list[(size_t)i * 3], c_list[(size_t)i * 3 + 1], c_list[(size_t)i * 3 + 2]};
        const T cj[3] = {c_list[(size_t)j * 3], c_list[(size_t)j * 3 + 1], c_list[(size_t)j * 3 + 2]};
        const T g[3] = {g6, g8, g10}, h[3] = {h6, h8, h10};
        T irp = ir6, e = 0, dEdr2 = 0, dm = 0;
        T gci[3] = {0, 0, 0}, gcj[3] = {0, 0, 0};
        const int np = (pmax - 4) / 2;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < np) {
                const T cc = ci[k] * cj[k], f = (m + g[k] - (T)1);
                e += f * cc * irp;
                // d/dr2: h kappa^2 cc irp - (p/2) f cc irp / r2
                dEdr2 += cc * irp * (h[k] * kappa * kappa - (T)(3 + k) * f * ir2);
                dm += cc * irp;
                gci[k] = f * cj[k] * irp; gcj[k] = f * ci[k] * irp;
            }
            irp *= ir2;
        }
        acc_e = (double)e;
        if (want_pg) {
            acc_ms[sidx] = (double)dm;
            if (dc != nullptr) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { atomicAdd(dc + (size_t)i * 3 + k, gci[k]); atomicAdd(dc + (size_t)j * 3 + k, gcj[k]); }
            }
        }
        if (want_grad) {
            T fv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                fv[k] = 2 * dEdr2 * d[k];
                atomicAdd(dpos + (size_t)i * 3 + k, fv[k]);
                atomicAdd(dpos + (size_t)j * 3 + k, -fv[k]);
            }
            if (want_vir) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) acc_box[3 * a + b] = -(double)(sh[a] * fv[b]);
            }
        }
    }
    double v1[1] = {acc_e};
    block_accumulate<1>(v1, red, scalars + ADMP_S_E_REAL);
    if (want_vir && want_grad) block_accumulate<9>(acc_box, red, scalars + ADMP_S_DBOX);
    if (want_pg) block_accumulate<5>(acc_ms, red, scalars + ADMP_S_DMSCALE);
}

// Tang-Toennies short-range kernel through the generic pair driver: admp/pairwise.py:45-113 (TT_damping_qq_c6_kernel), and
// its extension to the C8 / C10 terms with their own Tang-Toennies damping (pc8 / pc10 non-null):
//   f = 2625.5 a e^-br - 2625.5 e^-br (1 + br) q / br + sum_{n in 6[,8,10]} e^-br P_n(br) c_n,i c_n,j / r^n,  P_n = sum_{k<=n} x^k/k!
template <typename T>
__global__ void __launch_bounds__(128)
tt_pair_kernel(int64_t n_rows, int n_atoms, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos,
               const int32_t* __restrict__ pairs, const int32_t* __restrict__ cov_off, const int32_t* __restrict__ cov_idx,
               const int8_t* __restrict__ cov_nb, const T* __restrict__ mScales, const T* __restrict__ pa,
               const T* __restrict__ pb, const T* __restrict__ pq, const T* __restrict__ pc, const T* __restrict__ pc8,
               const T* __restrict__ pc10, uint32_t flags, T* __restrict__ dpos, T* __restrict__ dparams, double* __restrict__ scalars) {
    __shared__ double red[10 * 4];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc_e = 0.0, acc_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, acc_ms[5] = {0, 0, 0, 0, 0};
    const bool want_grad = (flags & ADMP_WANT_GRAD) != 0, want_vir = (flags & ADMP_WANT_VIRIAL) != 0,
               want_pg = (flags & ADMP_WANT_PGRAD) != 0;
    int i = 0, j = 0;
    bool live = false;
    if (p < n_rows) { i = pairs[2 * p]; j = pairs[2 * p + 1]; live = (i < j) && (i >= 0) && (j < n_atoms); }
    if (live) {
        T d[3] = {pos[3 * i] - pos[3 * j], pos[3 * i + 1] - pos[3 * j + 1], pos[3 * i + 2] - pos[3 * j + 2]};
        T sh[3];
        min_image(*Bp, d, sh);
        const T r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const T r = sqrt(r2);
        const int sidx = scale_index(cov_off, cov_idx, cov_nb, i, j);
        const T m = mScales[sidx];
        const T ai = pa[i], aj = pa[j], bi = pb[i], bj = pb[j], qi = pq[i], qj = pq[j];
        const T a = sqrt(ai * aj), b = sqrt(bi * bj), q = qi * qj;
        const T bohr = (T)1.889726878, ha = (T)2625.5;
        const T br = b * r * bohr, ex = exp(-br);
        // P_n and x^n/n! for n = 6, 8, 10 (d/dx [e^-x P_n] = -e^-x x^n/n!)
        T poly = 1, term = 1, P[3] = {0, 0, 0}, top[3] = {0, 0, 0};
#pragma unroll
        for (int k = 1; k <= 10; ++k) {
            term *= br / (T)k;
            poly += term;
            if (k == 6) { P[0] = poly; top[0] = term; }
            if (k == 8) { P[1] = poly; top[1] = term; }
            if (k == 10) { P[2] = poly; top[2] = term; }
        }
        const T ir2 = (T)1 / r2, ir6 = ir2 * ir2 * ir2;
        const T irn[3] = {ir6, ir6 * ir2, ir6 * ir2 * ir2};
        const T* pcs[3] = {pc, pc8, pc10};
        const T f1 = ha * a * ex;
        const T f2 = -ha * ex * ((T)1 + br) * q / br;
        T f3 = 0, d3 = 0, nf3 = 0;          // dispersion energy, its d/d(br), sum n * term_n
        T cn_i[3] = {0, 0, 0}, cn_j[3] = {0, 0, 0};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (pcs[t] != nullptr) {
                cn_i[t] = pcs[t][i]; cn_j[t] = pcs[t][j];
                const T cc = cn_i[t] * cn_j[t];
                const T e_t = ex * P[t] * cc * irn[t];
                f3 += e_t;
                d3 -= ex * top[t] * cc * irn[t];
                nf3 += (T)(6 + 2 * t) * e_t;
            }
        }
        const T e = (f1 + f2 + f3) * m;
        acc_e = (double)e;
        const T d1 = -f1;
        // d/dbr [ -e^{-br}(1+br)/br ] = e^{-br} (1 + (1+br)/br^2)
        const T d2 = ha * ex * q * ((T)1 + ((T)1 + br) / (br * br));
        const T dE_dbr = (d1 + d2 + d3) * m;
        const T dE_dr = dE_dbr * b * bohr - nf3 * m / r;
        if (want_grad) {
            T fv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                fv[k] = dE_dr * d[k] / r;
                atomicAdd(dpos + (size_t)i * 3 + k, fv[k]);
                atomicAdd(dpos + (size_t)j * 3 + k, -fv[k]);
            }
            if (want_vir) {
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) acc_box[3 * aa + bb] = -(double)(sh[aa] * fv[bb]);
            }
        }
        if (want_pg) {
            acc_ms[sidx] = (double)(f1 + f2 + f3);
            if (dparams != nullptr) {
                const size_t n = (size_t)n_atoms;
                // a = sqrt(ai aj): da/dai = a/(2 ai)
                const T ea = m * ha * ex;
                atomicAdd(dparams + i, ea * a / (2 * ai)); atomicAdd(dparams + j, ea * a / (2 * aj));
                const T eb = dE_dbr * r * bohr;      // dE/db
                atomicAdd(dparams + n + i, eb * b / (2 * bi)); atomicAdd(dparams + n + j, eb * b / (2 * bj));
                const T eq = m * (-ha * ex * ((T)1 + br) / br);
                atomicAdd(dparams + 2 * n + i, eq * qj); atomicAdd(dparams + 2 * n + j, eq * qi);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    if (pcs[t] != nullptr) {
                        const T ec = m * ex * P[t] * irn[t];
                        atomicAdd(dparams + (3 + t) * n + i, ec * cn_j[t]); atomicAdd(dparams + (3 + t) * n + j, ec * cn_i[t]);
                    }
                }
            }
        }
    }
    double v1[1] = {acc_e};
    block_accumulate<1>(v1, red, scalars + ADMP_S_E_REAL);
    if (want_vir && want_grad) block_accumulate<9>(acc_box, red, scalars + ADMP_S_DBOX);
    if (want_pg) block_accumulate<5>(acc_ms, red, scalars + ADMP_S_DMSCALE);
}

template <typename T>
void launch_disp_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, double kappa, int pmax, const void* pos,
                      const int32_t* pairs, const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb,
                      const void* c_list, const void* mS, uint32_t flags, void* dpos, void* dc, double* scalars) {
    if (n_rows <= 0) return;
    disp_pair_kernel<T><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(n_rows, n_atoms, B, (T)kappa, pmax, (const T*)pos, pairs, cov_off,
                                                                         cov_idx, cov_nb, (const T*)c_list, (const T*)mS, flags,
                                                                         (T*)dpos, (T*)dc, scalars);
}
template <typename T>
void launch_tt_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                    const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb, const void* mS, const void* a,
                    const void* b, const void* q, const void* c, const void* c8, const void* c10, uint32_t flags, void* dpos,
                    void* dparams, double* scalars) {
    if (n_rows <= 0) return;
    tt_pair_kernel<T><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(n_rows, n_atoms, B, (const T*)pos, pairs, cov_off, cov_idx, cov_nb,
                                                                       (const T*)mS, (const T*)a, (const T*)b, (const T*)q, (const T*)c,
                                                                       (const T*)c8, (const T*)c10, flags, (T*)dpos, (T*)dparams, scalars);
}
template void launch_disp_pair<double>(cudaStream_t, int64_t, int, const BoxInfo*, double, int, const void*, const int32_t*, const int32_t*,
                                       const int32_t*, const int8_t*, const void*, const void*, uint32_t, void*, void*, double*);
template void launch_disp_pair<float>(cudaStream_t, int64_t, int, const BoxInfo*, double, int, const void*, const int32_t*, const int32_t*,
                                      const int32_t*, const int8_t*, const void*, const void*, uint32_t, void*, void*, double*);
template void launch_tt_pair<double>(cudaStream_t, int64_t, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const int32_t*,
                                     const int8_t*, const void*, const void*, const void*, const void*, const void*, const void*, const void*, uint32_t, void*, void*, double*);
template void launch_tt_pair<float>(cudaStream_t, int64_t, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const int32_t*,
                                    const int8_t*, const void*, const void*, const void*, const void*, const void*, const void*, const void*, uint32_t, void*, void*, double*);

}  // namespace admp
