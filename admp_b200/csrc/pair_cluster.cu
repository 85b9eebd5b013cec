// Cluster pair kernel: the real-space multipole pair sum (admp/pme.py:479-729) over j-CLUSTER x i-LANE tiles.
//
// The flat kernel (pair.cu) evaluates one pair row per thread and needs 16 float64 reductions per pair END: on a
// liquid-density list those scattered reductions, not the arithmetic, bound it (profiles/r1o_ncu_full.md). Here a warp
// owns one j-cluster (up to 4 consecutive, covalently bonded atoms - a water molecule) and walks the UNION of the
// clusters' neighbour lists 32 atoms at a time: lane = one i atom, inner loop = the cluster's j atoms.
//   * j side: the 16 gradient components of a j atom are summed over the warp with a 16-value reduce-scatter and
//     kept in one register per (lane, j slot) across all chunks of the cluster: one reduction per component per
//     cluster instead of one per pair.
//   * i side: the lane's 16 components accumulate in registers over the cluster's j atoms (~2.7 pairs per lane and
//     chunk for water) and are flushed once per chunk.
//   * the j records sit in shared memory (uniform-address reads), the i records are gathered once per chunk.
// The pair SET is exactly the caller's: the tiles are built from the caller's rows (i < j, strictly sorted by one column then
// the other - the orders admp_nblist_build and jax_md's OrderedSparse produce; the sorted-by column becomes the cluster side) with a per-entry mask of which (i, j_k) pairs are
// listed and their scale index; no pair is added or dropped. Any other row order, or a sparse list (gas-like boxes:
// fewer than ~1 full chunk per cluster), keeps the flat kernel: a device-side flag selects which of the two kernels
// does the work, so there is no host synchronisation and both launches are graph-capturable.
#include "pair_math.cuh"

namespace admp {

// ------------------------------------------------------------------------------------------ tile construction
// state[0] = rows are NOT strictly sorted by (column 0, column 1); state[1] = live rows (they must form a prefix);
// state[2] = use the cluster kernel (decision); state[3] = group column gc (the cluster side), the lanes take the other;
// state[4] = rows are NOT strictly sorted by (column 1, column 0); state[5] = live rows do not form a prefix.
// admp_nblist_build and the oracle list rows by (i, j) ascending (gc = 0); jax_md's OrderedSparse groups by the receiver.
__global__ void __launch_bounds__(256)
cluster_check_kernel(int64_t n_rows, const int32_t* __restrict__ pairs, const int8_t* __restrict__ sidx, int32_t* __restrict__ state) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    if (p < n_rows) {
        live = sidx[p] >= 0;
        if (live && p > 0) {
            const int a = pairs[2 * p], b = pairs[2 * p + 1];
            const int ap = pairs[2 * p - 2], bp = pairs[2 * p - 1];
            if (sidx[p - 1] < 0) state[5] = 1;
            if (ap > a || (ap == a && bp >= b)) state[0] = 1;
            if (bp > b || (bp == b && ap >= a)) state[4] = 1;
        }
    }
    // the live rows must form a prefix (state[5] otherwise), so their number is the index after the last live row:
    // one plain store instead of a same-address atomic per warp
    if (live && (p + 1 == n_rows || sidx[p + 1] < 0)) state[1] = (int32_t)(p + 1);
}

__global__ void cluster_decide_kernel(int n_clusters, int min_rows_per_cluster, int force, int32_t* __restrict__ state) {
    // force: 0 = auto, 1 = cluster whenever the order allows, -1 = never
    int use = 0, gc = 0;
    const bool ok0 = state[0] == 0, ok1 = state[4] == 0;
    if (force >= 0 && state[5] == 0 && (ok0 || ok1) && n_clusters > 0 && state[1] > 0) {
        gc = ok0 ? 0 : 1;
        use = force > 0 ? 1 : ((int64_t)state[1] >= (int64_t)min_rows_per_cluster * n_clusters);
    }
    state[2] = use;
    state[3] = gc;
}

// first row of every atom in the group column: lower bound over the sorted live prefix (no row: the next atom's start)
__global__ void __launch_bounds__(256)
cluster_rowstart_kernel(int n_atoms, const int32_t* __restrict__ pairs, int32_t* __restrict__ row_start, const int32_t* __restrict__ state) {
    if (state[2] == 0) return;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n_atoms) return;
    const int gc = state[3];
    int lo = 0, hi = state[1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pairs[2 * (int64_t)mid + gc] < a) lo = mid + 1; else hi = mid;
    }
    row_start[a] = lo;
}

__global__ void __launch_bounds__(256)
cluster_build_kernel(int64_t n_rows, const int32_t* __restrict__ pairs, const int8_t* __restrict__ sidx,
                     const int32_t* __restrict__ cl_of, const int32_t* __restrict__ cl_first, const int32_t* __restrict__ row_start,
                     int32_t* __restrict__ ent_i, uint32_t* __restrict__ ent_m, int32_t* __restrict__ cl_extra,
                     const int32_t* __restrict__ state) {
    if (state[2] == 0) return;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rows) return;
    const int s = sidx[p];
    if (s < 0) return;
    const int gc = state[3], oc = 1 - gc;
    const int i = pairs[2 * p + oc], j = pairs[2 * p + gc];          // j: cluster side, i: lane side
    const int c = cl_of[j], first = cl_first[c], k = j - first;
    const uint32_t bits = (uint32_t)(1 | (s << 1)) << (4 * k);       // 4 bits per slot: listed, scale index (3 bits)
    const int base = row_start[first], n0 = row_start[first + 1] - base;
    if (k == 0) {
        ent_i[p] = i;
        atomicOr(&ent_m[p], bits);
        return;
    }
    // is i already an entry of slot 0's (sorted) list?
    int lo = 0, hi = n0;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pairs[2 * (int64_t)(base + mid) + oc] < i) lo = mid + 1; else hi = mid;
    }
    if (lo < n0 && pairs[2 * (int64_t)(base + lo) + oc] == i) {
        atomicOr(&ent_m[base + lo], bits);
    } else {
        const int e = atomicAdd(&cl_extra[c], 1);
        ent_i[base + n0 + e] = i;
        atomicOr(&ent_m[base + n0 + e], bits);
    }
}

// ------------------------------------------------------------------------------------------ the pair arithmetic
// One pair (I = the lane's atom, J = a cluster atom read from shared memory). ir / jr: packed records
// (0-2 position, 3-12 Cartesian multipoles, 13-15 induced dipole, 16 polarizability, 17 Thole width).
// Accumulates the I-side gradient into gi (0-2 dE/dr, 3-12 dE/dM, 13-15 dE/dU) and RETURNS the J side in gj.
template <typename T, bool POL, int MODE, bool PG>
__device__ __forceinline__ void cluster_pair(const BoxInfo& B, T kappa, const T (&ir)[20], const T* __restrict__ jr, T mscale, T pscale, T w0,
                                             bool want_grad, bool want_vir, T (&gi)[16], T (&gj)[16], double& acc_e,
                                             double (&acc_box)[9], T& dm_out, T& dp_out, T& eth_out, T& edpi_out, T& edpj_out) {
    T mj[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) mj[k] = jr[3 + k];
    const T* mi = ir + 3;
    T d[3] = {ir[0] - jr[0], ir[1] - jr[1], ir[2] - jr[2]};
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
    T uI[3] = {0, 0, 0}, uJ[3] = {0, 0, 0}, pI = 0, pJ = 0, polI = 0, polJ = 0;
    IndCoef<T> C;
    if (POL) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { uI[k] = ir[13 + k]; uJ[k] = jr[13 + k]; }
        pI = dot3(uI, n); pJ = dot3(uJ, n);
        polI = ir[16]; polJ = jr[16];
        ind_coeffs<T, MODE == 0, true>(R, pscale, w0, ir[17], jr[17], ir[18], jr[18], C);
    }
    if (MODE == 1) {
        const T B1 = C.B[0], B2 = C.B[1], B3 = C.B[2], B5 = C.B[3], B6 = C.B[4], C2 = C.B[5], C3 = C.B[6];
        const T e_pI = -B1 * qJ + B2 * dJ - B5 * tJ + C2 * pJ;
        const T e_pJ = B1 * qI + B2 * dI + B5 * tI + C2 * pI;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            gi[13 + k] += e_pI * n[k] + B3 * muJ[k] - B6 * vJ[k] + C3 * uJ[k];
            gj[13 + k] = e_pJ * n[k] + B3 * muI[k] + B6 * vI[k] + C3 * uI[k];
        }
        return;
    }
    T A[10], dA[10], mA[10];
    perm_coeffs<T, true>(R, mscale, A, dA, mA);
    const T mm = dot3(muI, muJ), gJI = dot3(muJ, vI), gIJ = dot3(muI, vJ), vv = dot3(vI, vJ);
    const T TT = TI[0] * TJ[0] + TI[3] * TJ[3] + TI[5] * TJ[5] + 2 * (TI[1] * TJ[1] + TI[2] * TJ[2] + TI[4] * TJ[4]);
    const T inv[10] = {qI * qJ, qI * dJ - dI * qJ, dI * dJ, mm, tI * qJ + qI * tJ, tI * dJ - dI * tJ, gJI - gIJ, tI * tJ, vv, TT};
    T e = 0, dEdr = 0, dm = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) { e += A[k] * inv[k]; dEdr += dA[k] * inv[k]; if (PG) dm += mA[k] * inv[k]; }
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
    if (PG) {
        dm_out = dm;
        if (POL) {
            T dp = 0, da = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) { dp += C.pB[k] * inv2[k]; da += C.aB[k] * inv2[k]; }
            dp_out = dp;
            eth_out = da * C.au_a * C.da_dth;
            if (!C.trimmed) {
                const T e_dmp = da * C.au_d * C.dmp * (T)(1.0 / 6);
                edpi_out = e_dmp / polI; edpj_out = e_dmp / polJ;
            }
        }
    }
    if (!want_grad) return;
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
        gi[k] += fv[k];
        gj[k] = -fv[k];
    }
    if (want_vir) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) acc_box[3 * a + b] -= (double)(sh[a] * fv[b]);
    }
    gi[3] += g_qI;
    gj[3] = g_qJ;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gi[4 + k] += e_dI * n[k] + A[3] * muJ[k] - A[6] * vJ[k] + B3 * uJ[k];
        gj[4 + k] = e_dJ * n[k] + A[3] * muI[k] + A[6] * vI[k] + B3 * uI[k];
    }
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
        gi[7 + k] += e_tI * n[a] * n[b] * mult + ((a == b) ? wI[a] * n[a] : wI[a] * n[b] + wI[b] * n[a]) + A[9] * TJ[k] * mult;
        gj[7 + k] = e_tJ * n[a] * n[b] * mult + ((a == b) ? wJ[a] * n[a] : wJ[a] * n[b] + wJ[b] * n[a]) + A[9] * TI[k] * mult;
    }
    if (POL) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            gi[13 + k] += e_pI * n[k] + B3 * muJ[k] - B6 * vJ[k] + C3 * uJ[k];
            gj[13 + k] = e_pJ * n[k] + B3 * muI[k] + B6 * vI[k] + C3 * uI[k];
        }
    }
}

constexpr int CL_MAX = 4;            // atoms per j-cluster
constexpr int CL_WARPS = 4;          // warps per block (each owns whole clusters)

template <typename T, bool POL, int MODE, bool PG>
__global__ void __launch_bounds__(32 * CL_WARPS, (MODE == 1 || sizeof(T) == 4) ? 4 : 2)
pme_cluster_kernel(int n_clusters, const BoxInfo* __restrict__ Bp, T kappa, const int32_t* __restrict__ cl_first,
                   const int32_t* __restrict__ cl_size, const int32_t* __restrict__ row_start, const int32_t* __restrict__ cl_extra,
                   const int32_t* __restrict__ ent_i, const uint32_t* __restrict__ ent_m, const T* __restrict__ rec,
                   const T* __restrict__ Ucur, const T* __restrict__ mScales, const T* __restrict__ pScales, uint32_t flags,
                   T* __restrict__ dpos, T* __restrict__ G, T* __restrict__ F, T* __restrict__ dpol, T* __restrict__ dth,
                   double* __restrict__ scalars, const int32_t* __restrict__ state) {
    if (state[2] == 0) return;                                   // the flat kernel does the work
    constexpr int REC = PairRec<T>::STRIDE, EPC = PairRec<T>::EPC, NCH = PairRec<T>::chunks_all();
    __shared__ double red[10 * CL_WARPS];
    __shared__ BoxInfo sB;
    __shared__ T sScale[10], sW0[5];
    __shared__ __align__(16) T sJ[CL_WARPS][CL_MAX][NCH * EPC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < (int)(sizeof(BoxInfo) / sizeof(double))) reinterpret_cast<double*>(&sB)[tid] = reinterpret_cast<const double*>(Bp)[tid];
    if (tid < 5) { sScale[tid] = mScales[tid]; sScale[5 + tid] = POL ? pScales[tid] : (T)0; sW0[tid] = POL ? thole_switch_w0<T>(pScales[tid]) : (T)0; }
    __syncthreads();
    const BoxInfo& B = sB;
    const bool want_grad = MODE == 1 || (flags & ADMP_WANT_GRAD) != 0, want_vir = (flags & ADMP_WANT_VIRIAL) != 0;
    double acc_e = 0.0;
    double acc_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double acc_ms[5] = {0, 0, 0, 0, 0}, acc_ps[5] = {0, 0, 0, 0, 0};
    const int gw = blockIdx.x * CL_WARPS + warp, nw = gridDim.x * CL_WARPS;
    for (int c = gw; c < n_clusters; c += nw) {
        const int first = cl_first[c], csize = cl_size[c];
        const int base = row_start[first];
        const int cnt = (row_start[first + 1] - base) + cl_extra[c];
        if (cnt == 0) continue;
        // the cluster's j records -> shared memory (16-byte chunks, one per lane)
        __syncwarp();
        for (int q = lane; q < csize * NCH; q += 32) {
            const int k = q / NCH, ch = q - k * NCH;
            const PairChunk<T> v = *(reinterpret_cast<const PairChunk<T>*>(rec + (size_t)(first + k) * REC) + ch);
#pragma unroll
            for (int e = 0; e < EPC; ++e) sJ[warp][k][ch * EPC + e] = v.v[e];
        }
        if (MODE == 1 && lane < 3 * csize) sJ[warp][lane / 3][13 + lane % 3] = Ucur[(size_t)first * 3 + lane];   // this cycle's U
        __syncwarp();
        T jacc[CL_MAX];              // lane L holds component (L >> 1) & 15 of slot k (MODE 0); F component L of slot k (MODE 1, L < 3)
#pragma unroll
        for (int k = 0; k < CL_MAX; ++k) jacc[k] = (T)0;
        T jth[CL_MAX], jpol[CL_MAX];
#pragma unroll
        for (int k = 0; k < CL_MAX; ++k) { jth[k] = (T)0; jpol[k] = (T)0; }
        for (int e0 = 0; e0 < cnt; e0 += 32) {
            const int e = e0 + lane;
            const bool act = e < cnt;
            const int i = act ? ent_i[base + e] : first;
            const uint32_t m = act ? ent_m[base + e] : 0u;
            T ir[20];
            {
                const PairChunk<T>* src = reinterpret_cast<const PairChunk<T>*>(rec + (size_t)i * REC);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const PairChunk<T> v = src[ch];
#pragma unroll
                    for (int q = 0; q < EPC; ++q)
                        ir[ch * EPC + q] = v.v[q];
                }
                if (MODE == 1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) ir[13 + k] = Ucur[(size_t)i * 3 + k];
                }
            }
            T gi[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) gi[k] = (T)0;
            T ith = 0, ipol = 0;
            // one copy of the pair arithmetic (~3 000 instructions): the slot loop must stay rolled or the kernel outgrows
            // the instruction cache (4 copies = 130 KB of code: "no instruction" became the top stall)
#pragma unroll 1
            for (int k = 0; k < csize; ++k) {
                const uint32_t bits = (m >> (4 * k)) & 15u;
                const bool on = bits & 1u;
                if (!__any_sync(0xffffffffu, on)) continue;
                T gj[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) gj[q] = (T)0;
                T dm = 0, dp = 0, eth = 0, edpi = 0, edpj = 0;
                if (on) {
                    const int s = (int)(bits >> 1);
                    cluster_pair<T, POL, MODE, PG>(B, kappa, ir, sJ[warp][k], sScale[s], sScale[5 + s], sW0[s], want_grad, want_vir, gi, gj,
                                                   acc_e, acc_box, dm, dp, eth, edpi, edpj);
                    if (PG) {
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            acc_ms[q] += (q == s) ? (double)dm : 0.0;
                            acc_ps[q] += (q == s) ? (double)dp : 0.0;
                        }
                        ith += eth; ipol += edpi;
                    }
                }
                T r = (T)0;
                if (MODE == 1) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const T s3 = warp_sum(gj[13 + q]);
                        if (lane == q) r = s3;
                    }
                } else if (want_grad) {
                    r = warp_reduce_scatter16(gj, lane);
                }
#pragma unroll
                for (int q = 0; q < CL_MAX; ++q) {
                    jacc[q] += (q == k) ? r : (T)0;
                    if (PG && POL) { jth[q] += (q == k) ? eth : (T)0; jpol[q] += (q == k) ? edpj : (T)0; }
                }
            }
            // i side: one flush per chunk
            if (m != 0u) {
                if (MODE == 0 && want_grad) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) atomicAdd(dpos + (size_t)i * 3 + q, gi[q]);
#pragma unroll
                    for (int q = 0; q < 10; ++q) atomicAdd(G + (size_t)i * 10 + q, gi[3 + q]);
                }
                if (POL && want_grad) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) atomicAdd(F + (size_t)i * 3 + q, gi[13 + q]);
                }
                if (PG && POL) {
                    if (dth != nullptr) atomicAdd(dth + i, ith);
                    if (dpol != nullptr && ipol != (T)0) atomicAdd(dpol + i, ipol);
                }
            }
        }
        // j side: one flush per cluster
#pragma unroll
        for (int k = 0; k < CL_MAX; ++k) {
            if (k < csize) {
                const size_t j = (size_t)(first + k);
                if (MODE == 1) {
                    if (lane < 3) atomicAdd(F + j * 3 + lane, jacc[k]);
                } else if (want_grad) {
                    const int v = (lane >> 1) & 15;
                    if ((lane & 1) == 0) {
                        if (v < 3) atomicAdd(dpos + j * 3 + v, jacc[k]);
                        else if (v < 13) atomicAdd(G + j * 10 + (v - 3), jacc[k]);
                        else if (POL) atomicAdd(F + j * 3 + (v - 13), jacc[k]);
                    }
                }
                if (PG && POL) {
                    const T st = warp_sum(jth[k]), sp = warp_sum(jpol[k]);
                    if (lane == 0) {
                        if (dth != nullptr) atomicAdd(dth + j, st);
                        if (dpol != nullptr && sp != (T)0) atomicAdd(dpol + j, sp);
                    }
                }
            }
        }
    }
    if (MODE == 0) {
        double v1[1] = {acc_e};
        block_accumulate<1>(v1, red, scalars + ADMP_S_E_REAL);
        if (want_vir && want_grad) block_accumulate<9>(acc_box, red, scalars + ADMP_S_DBOX);
        if (PG) {
            block_accumulate<5>(acc_ms, red, scalars + ADMP_S_DMSCALE);
            if (POL) block_accumulate<5>(acc_ps, red, scalars + ADMP_S_DPSCALE);
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
void launch_cluster_prepare(cudaStream_t st, int64_t n_rows, int n_atoms, int n_clusters, const int32_t* pairs, const int8_t* sidx,
                            const ClusterWork& w, int force) {
    cudaMemsetAsync(w.state, 0, sizeof(int32_t) * 8, st);
    if (n_rows <= 0 || n_clusters <= 0 || force < 0) return;
    cudaMemsetAsync(w.ent_m, 0, sizeof(uint32_t) * (size_t)n_rows, st);
    cudaMemsetAsync(w.cl_extra, 0, sizeof(int32_t) * (size_t)n_clusters, st);
    const unsigned g = (unsigned)((n_rows + 255) / 256);
    cluster_check_kernel<<<g, 256, 0, st>>>(n_rows, pairs, sidx, w.state);
    cluster_decide_kernel<<<1, 1, 0, st>>>(n_clusters, w.min_rows_per_cluster, force, w.state);
    cluster_rowstart_kernel<<<(unsigned)((n_atoms + 256) / 256), 256, 0, st>>>(n_atoms, pairs, w.row_start, w.state);
    cluster_build_kernel<<<g, 256, 0, st>>>(n_rows, pairs, sidx, w.cl_of, w.cl_first, w.row_start, w.ent_i, w.ent_m, w.cl_extra, w.state);
}

template <typename T, bool POL, int MODE, bool PG>
static void launch_cluster_t(cudaStream_t st, int n_clusters, const BoxInfo* B, double kappa, const ClusterWork& w, const void* rec,
                             const void* U, const void* mS, const void* pS, uint32_t flags, void* dpos, void* G, void* F, void* dpol,
                             void* dth, double* scalars) {
    static int grid_cap = 0;
    auto kern = pme_cluster_kernel<T, POL, MODE, PG>;
    if (grid_cap == 0) {
        int occ = 1, dev = 0, nsm = 148;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * CL_WARPS, 0);
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        grid_cap = nsm * (occ > 0 ? occ : 1);
    }
    const int need = (n_clusters + CL_WARPS - 1) / CL_WARPS;
    const unsigned grid = (unsigned)(need < grid_cap ? need : grid_cap);
    kern<<<grid, 32 * CL_WARPS, 0, st>>>(n_clusters, B, (T)kappa, w.cl_first, w.cl_size, w.row_start, w.cl_extra, w.ent_i, w.ent_m,
                                         (const T*)rec, (const T*)U, (const T*)mS, (const T*)pS, flags, (T*)dpos, (T*)G, (T*)F, (T*)dpol,
                                         (T*)dth, scalars, w.state);
}

template <typename T>
void launch_pme_cluster(cudaStream_t st, int n_clusters, const BoxInfo* B, double kappa, const ClusterWork& w, const void* rec,
                        const void* U, const void* mS, const void* pS, int mode, uint32_t flags, void* dpos, void* G, void* F,
                        void* dpol, void* dth, double* scalars) {
    if (n_clusters <= 0) return;
    const bool polz = (U != nullptr);
    const bool pg = (flags & ADMP_WANT_PGRAD) != 0;
#define ADMP_CL_ARGS st, n_clusters, B, kappa, w, rec, U, mS, pS, flags, dpos, G, F, dpol, dth, scalars
    if (mode == 1) {
        if (polz) launch_cluster_t<T, true, 1, false>(ADMP_CL_ARGS);
    } else if (polz) {
        if (pg) launch_cluster_t<T, true, 0, true>(ADMP_CL_ARGS); else launch_cluster_t<T, true, 0, false>(ADMP_CL_ARGS);
    } else {
        if (pg) launch_cluster_t<T, false, 0, true>(ADMP_CL_ARGS); else launch_cluster_t<T, false, 0, false>(ADMP_CL_ARGS);
    }
#undef ADMP_CL_ARGS
}
template void launch_pme_cluster<double>(cudaStream_t, int, const BoxInfo*, double, const ClusterWork&, const void*, const void*,
                                         const void*, const void*, int, uint32_t, void*, void*, void*, void*, void*, double*);
template void launch_pme_cluster<float>(cudaStream_t, int, const BoxInfo*, double, const ClusterWork&, const void*, const void*,
                                        const void*, const void*, int, uint32_t, void*, void*, void*, void*, void*, double*);

}  // namespace admp
