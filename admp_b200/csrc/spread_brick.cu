// Brick-staged B-spline multipole spread: the mesh is written ONCE, with coalesced stores, and the zero-fill is part of it.
// Replaces admp/recip.py:313-392 (Q = zeros; Q.at[...].add of the 216 stencil values per atom) for single-GPU meshes;
// recip.cu keeps the one-warp-per-atom scatter (global atomics on a zero-filled mesh) for x-slab decomposed meshes,
// atom sub-ranges and meshes too small for bricks.
//
// Layout: the mesh is cut into bricks of 16 x 16 x BZ points (BZ = 16 or 32, z fastest; edge bricks are partial). Once per
// evaluation the atoms are binned by the brick that holds their stencil anchor (count -> scan -> fill: `anchor` holds
// (i0, j0, k0, atom) sorted by home brick); all 31 spreads of a polarizable evaluation reuse the bins. One block owns one
// brick: it walks the atoms of the <= 27 (normally 8) home bricks whose stencils can reach it, keeps those whose 6^3
// footprint intersects the brick (integer tests on the stored anchors), loads their precomputed records (brick_prep_kernel:
// splines and y-contracted multipole coefficients, evaluated once per atom and spread) 32 atoms at a time, and accumulates
// into a shared-memory tile: half-warp h owns x plane h of the tile, its 16 lanes cover the 36 (y, z) points of an atom's
// stencil plane in 3 passes - no two half-warps ever touch the same address, so there are no atomics at all. Finally the
// tile is streamed to global memory in full 128/256-byte rows: every mesh point written exactly once, no separate zero-fill.
// Algorithmic bytes: w*G (one write of the mesh) + Na*(3 + n_comp)*w.
// STATUS: opt-in (admp_ctx_set_spread / ADMP_SPREAD=bricks). Measured on B200 it loses to zero-fill + per-atom scatter on every
// workload (308x616x616, 98 304 atoms: 0.55 vs 0.26 ms; 154^3: 0.034 vs 0.020 ms; dense 308^3, 786 432 atoms: 1.9 vs 0.93 ms):
// each atom is visited by ~2.26 bricks and a visit costs as many instructions as the whole per-atom scatter, so the kernel is
// issue-bound (ncu: 305 M warp instructions per launch at 308x616x616, 29 % IMAD, IPC 2.0, DRAM 21 %), while the scatter runs
// at the L2 atomic rate (~170-190 G RED.64/s) and the zero-fill at the HBM peak (profiles/r2i_brick_spread.md).
#include <cstdlib>

#include "kernels.h"
#include "bspline.cuh"

namespace admp {

constexpr int BRX = 16, BRY = 16;     // brick footprint in x and y (mesh points)
constexpr int BR_THREADS = 256, BR_WARPS = BR_THREADS / 32;
constexpr int BR_NB = 32;             // atoms staged per batch (3 spline threads + 1 coefficient thread each)
constexpr int BR_CHUNK = BR_THREADS;  // candidate atoms filtered per round

// ---------------------------------------------------------------------------- binning (once per evaluation)
template <typename T>
__global__ void __launch_bounds__(128)
brick_count_kernel(int n, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos, BrickGeom g, int4* __restrict__ tmp,
                   int32_t* __restrict__ count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const BoxInfo& B = *Bp;
    const double rx = (double)pos[3 * a], ry = (double)pos[3 * a + 1], rz = (double)pos[3 * a + 2];
    double f;
    int i0, j0, k0;
    mesh_anchor(B, rx, ry, rz, 0, f, i0);
    mesh_anchor(B, rx, ry, rz, 1, f, j0);
    mesh_anchor(B, rx, ry, rz, 2, f, k0);
    const int id = ((i0 / BRX) * g.nb[1] + j0 / BRY) * g.nb[2] + k0 / g.bz;
    tmp[a] = make_int4(i0, j0, k0, id);
    atomicAdd(count + id, 1);
}

// exclusive scan of the brick histogram by one block (<= ~1e5 bricks; once per evaluation); leaves count = 0 (fill cursors)
__global__ void __launch_bounds__(1024)
brick_scan_kernel(int nb, int32_t* __restrict__ count, int32_t* __restrict__ start) {
    __shared__ int wtot[32], wpre[32];
    __shared__ int tot_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wtot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            const int t = wtot[lane];
            int s = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            wpre[lane] = s - t;                 // exclusive prefix of the warp totals
            if (lane == 31) tot_s = s;
        }
        __syncthreads();
        if (i < nb) {
            start[i] = carry + wpre[warp] + x - v;
            count[i] = 0;
        }
        carry += tot_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[nb] = carry;
}

__global__ void __launch_bounds__(128)
brick_fill_kernel(int n, const int4* __restrict__ tmp, const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                  int4* __restrict__ anchor) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int4 t = tmp[a];
    const int slot = start[t.w] + atomicAdd(cursor + t.w, 1);
    anchor[slot] = make_int4(t.x, t.y, t.z, a);
}

// ---------------------------------------------------------------------------- per-atom records (once per spread)
// Everything a brick needs from an atom, evaluated ONCE per atom (a brick-side evaluation would repeat it for each of the
// ~2 bricks an atom reaches and put the spline recursion on every brick's critical path). Record of sorted slot s:
//   multipoles (72 reals): wx[18] = M6, M6', M6'' at the 6 x offsets | sp[6][6] (per y offset ib: the y-contracted
//     coefficients P0 P1 P2 Q0 Q1 R0 of the three z-derivative orders) | wz[18]
//       t0(ia, ib) = a0 P0 + a1 P1 + a2 P2 ; t1 = a0 Q0 + a1 Q1 ; t2 = a0 R0 ; value(ia, ib, ic) = t0 c0 + t1 c1 + t2 c2
//   charges only (18 reals): wx[6] | q * wy[6] | wz[6]
template <bool MULTIPOLE> struct BrickRec { static constexpr int RS = MULTIPOLE ? 72 : 18; };
constexpr int PREP_ATOMS = 32, PREP_THREADS = 128;

template <typename T, bool MULTIPOLE>
__global__ void __launch_bounds__(PREP_THREADS)
brick_prep_kernel(int n, const BoxInfo* __restrict__ Bp, const int4* __restrict__ anchor, const T* __restrict__ pos,
                  const T* __restrict__ M, int m_stride, const T* __restrict__ U, T* __restrict__ rec) {
    constexpr int NW = MULTIPOLE ? 18 : 6, RS = BrickRec<MULTIPOLE>::RS;
    __shared__ T sw[PREP_ATOMS * 3 * NW];
    __shared__ T sc[PREP_ATOMS * 10];
    __shared__ T so[PREP_ATOMS * RS];
    const BoxInfo& B = *Bp;
    const int tid = threadIdx.x, s0 = blockIdx.x * PREP_ATOMS;
    const int nat = min(PREP_ATOMS, n - s0);
    if (tid < 3 * PREP_ATOMS) {
        const int at = tid / 3, d = tid - 3 * at;
        if (at < nat) {
            const int a = anchor[s0 + at].w;
            double f;
            int i0;
            mesh_anchor(B, (double)pos[3 * a], (double)pos[3 * a + 1], (double)pos[3 * a + 2], d, f, i0);
            bspline6<T, MULTIPOLE ? 3 : 1>((T)f, sw + (at * 3 + d) * NW);
        }
    } else {
        const int at = tid - 3 * PREP_ATOMS;
        if (at < nat) {
            const int a = anchor[s0 + at].w;
            const T* mrow = M + (size_t)a * m_stride;
            T* o = sc + at * 10;
            o[0] = mrow[0];
            if (MULTIPOLE) {
                T mu[3] = {mrow[1], mrow[2], mrow[3]};
                if (U != nullptr) { mu[0] += U[3 * a]; mu[1] += U[3 * a + 1]; mu[2] += U[3 * a + 2]; }
                const T Tm[9] = {mrow[4], mrow[5], mrow[6], mrow[5], mrow[7], mrow[8], mrow[6], mrow[8], mrow[9]};
                T N[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) N[k] = (T)B.nstar[k];
#pragma unroll
                for (int d = 0; d < 3; ++d) o[1 + d] = -(N[3 * d] * mu[0] + N[3 * d + 1] * mu[1] + N[3 * d + 2] * mu[2]);
                T NT[9];
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int b = 0; b < 3; ++b) NT[3 * d + b] = N[3 * d] * Tm[b] + N[3 * d + 1] * Tm[3 + b] + N[3 * d + 2] * Tm[6 + b];
                auto tf = [&](int d, int e) {
                    return (NT[3 * d] * N[3 * e] + NT[3 * d + 1] * N[3 * e + 1] + NT[3 * d + 2] * N[3 * e + 2]) * (T)(1.0 / 3);
                };
                o[4] = tf(0, 0); o[5] = 2 * tf(0, 1); o[6] = 2 * tf(0, 2); o[7] = tf(1, 1); o[8] = 2 * tf(1, 2); o[9] = tf(2, 2);
            }
        }
    }
    __syncthreads();
    if (MULTIPOLE) {
        for (int i = tid; i < nat * 6; i += PREP_THREADS) {
            const int at = i / 6, ib = i - 6 * at;
            const T* wy = sw + (at * 3 + 1) * NW;
            const T* q = sc + at * 10;            // q, muf[3], Tf[6] = (00, 2*01, 2*02, 11, 2*12, 22)
            const T b0 = wy[ib], b1 = wy[6 + ib], b2 = wy[12 + ib];
            T* o = so + at * RS + 18 + ib * 6;
            o[0] = q[0] * b0 + q[2] * b1 + q[7] * b2;
            o[1] = q[1] * b0 + q[5] * b1;
            o[2] = q[4] * b0;
            o[3] = q[3] * b0 + q[8] * b1;
            o[4] = q[6] * b0;
            o[5] = q[9] * b0;
        }
        for (int i = tid; i < nat * 36; i += PREP_THREADS) {
            const int at = i / 36, k = i - 36 * at;
            so[at * RS + (k < 18 ? k : 36 + k)] = sw[(at * 3 + (k < 18 ? 0 : 2)) * NW + (k < 18 ? k : k - 18)];
        }
    } else {
        for (int i = tid; i < nat * 18; i += PREP_THREADS) {
            const int at = i / 18, k = i - 18 * at, d = k / 6;
            const T v = sw[(at * 3 + d) * NW + (k - 6 * d)];
            so[at * RS + k] = (d == 1) ? v * sc[at * 10] : v;
        }
    }
    __syncthreads();
    T* out = rec + (size_t)s0 * RS;
    for (int i = tid; i < nat * RS; i += PREP_THREADS) out[i] = so[i];
}

// ---------------------------------------------------------------------------- spread
template <typename T, bool MULTIPOLE, int BZ>
struct BrickSmem {
    static constexpr int BZP = BZ + 6;                  // row stride = 6 mod 16: the 36 (ib, ic) offsets of a stencil plane
                                                        // are consecutive modulo the bank count (no conflicts within a pass)
    static constexpr int RS = BrickRec<MULTIPOLE>::RS;
    static constexpr size_t tile = (size_t)BRX * BRY * BZP;
    static constexpr size_t srec = (size_t)BR_NB * RS;
    static constexpr size_t bytes = (tile + srec) * sizeof(T) + (size_t)BR_CHUNK * sizeof(int2);
};

template <typename T, bool MULTIPOLE, int BZ>
__global__ void __launch_bounds__(BR_THREADS, 2)
spread_brick_kernel(const BoxInfo* __restrict__ Bp, BrickGeom g, const int32_t* __restrict__ start,
                    const int4* __restrict__ anchor, const T* __restrict__ rec, T* __restrict__ mesh) {
    using S = BrickSmem<T, MULTIPOLE, BZ>;
    constexpr int BZP = S::BZP, RS = S::RS;
    extern __shared__ __align__(16) unsigned char brick_smem[];
    T* tile = reinterpret_cast<T*>(brick_smem);
    T* srec = tile + S::tile;
    int2* L = reinterpret_cast<int2*>(srec + S::srec);   // (sorted slot, packed relative anchor) of the atoms that reach the brick
    __shared__ int cand_start[27], cand_pref[28];
    __shared__ int nL_s;

    const BoxInfo& B = *Bp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K1 = B.K[0], K2 = B.K[1], K3 = B.K[2];
    const int bzc = blockIdx.x % g.nb[2];
    const int byc = (blockIdx.x / g.nb[2]) % g.nb[1];
    const int bxc = blockIdx.x / (g.nb[2] * g.nb[1]);
    const int x0 = bxc * BRX, y0 = byc * BRY, z0 = bzc * BZ;
    const int bw = min(BRX, K1 - x0), bh = min(BRY, K2 - y0), bd = min(BZ, K3 - z0);

    // home bricks whose atoms can reach this brick: anchors in the cyclic interval [x0 - 5, x0 + bw - 1] per dimension
    const int fbx = ((x0 - 5 + K1) % K1) / BRX, ncx = (bxc - fbx + g.nb[0]) % g.nb[0] + 1;
    const int fby = ((y0 - 5 + K2) % K2) / BRY, ncy = (byc - fby + g.nb[1]) % g.nb[1] + 1;
    const int fbz = ((z0 - 5 + K3) % K3) / BZ, ncz = (bzc - fbz + g.nb[2]) % g.nb[2] + 1;
    const int nc = ncx * ncy * ncz;                      // <= 27 (three per dimension only next to a narrow edge brick)
    if (warp == 0) {
        int cnt = 0;
        if (lane < nc) {
            const int cz = lane % ncz, cy = (lane / ncz) % ncy, cx = lane / (ncz * ncy);
            const int id = (((fbx + cx) % g.nb[0]) * g.nb[1] + (fby + cy) % g.nb[1]) * g.nb[2] + (fbz + cz) % g.nb[2];
            const int s = start[id];
            cand_start[lane] = s;
            cnt = start[id + 1] - s;
        }
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane < 27) cand_pref[lane + 1] = x;
        if (lane == 0) cand_pref[0] = 0;
    }
    __syncthreads();
    const int C = cand_pref[nc];

    constexpr int RPW = 32 / BZ;                         // mesh rows (x, y) one warp store instruction covers
    const int zl = lane % BZ;
    if (C == 0) {                                        // nothing reaches this brick: it is part of the zero-fill
        for (int r0 = warp * RPW; r0 < BRX * BRY; r0 += BR_WARPS * RPW) {
            const int r = r0 + lane / BZ, x = r / BRY, y = r - x * BRY;
            if (x < bw && y < bh && zl < bd) mesh[((size_t)(x0 + x) * K2 + (y0 + y)) * K3 + z0 + zl] = (T)0;
        }
        return;
    }
    for (int k = tid; k < (int)S::tile; k += BR_THREADS) tile[k] = (T)0;

    // accumulation: half-warp hw owns x plane hw of the tile (no two half-warps ever touch the same address: no atomics);
    // its 16 lanes cover the 36 (ib, ic) points of an atom's stencil plane in 3 passes (pt = hl + 16 p)
    const int hw = tid >> 4, hl = tid & 15;
    int pib[3], pic[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const int pt = hl + 16 * p;
        pib[p] = pt / 6;
        pic[p] = pt - 6 * pib[p];
    }
    T* tplane = tile + hw * BRY * BZP;

    for (int base = 0; base < C; base += BR_CHUNK) {
        if (tid == 0) nL_s = 0;
        __syncthreads();
        // ---- filter: keep the candidates whose footprint intersects the brick (integer tests on the stored anchors)
        const int c = base + tid;
        bool ok = false;
        int2 ent = make_int2(0, 0);
        if (c < C) {
            int s = 0;
            while (c >= cand_pref[s + 1]) ++s;
            const int slot = cand_start[s] + (c - cand_pref[s]);
            const int4 an = anchor[slot];
            int rx = an.x - x0, ry = an.y - y0, rz = an.z - z0;
            if (rx > bw - 1) rx -= K1; else if (rx < -5) rx += K1;
            if (ry > bh - 1) ry -= K2; else if (ry < -5) ry += K2;
            if (rz > bd - 1) rz -= K3; else if (rz < -5) rz += K3;
            ok = rx >= -5 && rx <= bw - 1 && ry >= -5 && ry <= bh - 1 && rz >= -5 && rz <= bd - 1;
            ent = make_int2(slot, (rx + 5) | ((ry + 5) << 8) | ((rz + 5) << 16));
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&nL_s, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (ok) L[wbase + __popc(m & ((1u << lane) - 1u))] = ent;
        __syncthreads();
        const int nL = nL_s;

        for (int b0 = 0; b0 < nL; b0 += BR_NB) {
            const int nbat = min(BR_NB, nL - b0);
            // ---- the batch's records: independent, coalesced loads (one latency)
            for (int i = tid; i < nbat * RS; i += BR_THREADS) {
                const int at = i / RS, k = i - at * RS;
                srec[i] = rec[(size_t)L[b0 + at].x * RS + k];
            }
            __syncthreads();
            {
                for (int at = 0; at < nbat; ++at) {
                    const int e = L[b0 + at].y;
                    const int ia = hw - ((e & 255) - 5);
                    if ((unsigned)ia < 6u && hw < bw) {
                        const int ry = ((e >> 8) & 255) - 5, rz = (e >> 16) - 5;
                        const T* r = srec + at * RS;
                        if (MULTIPOLE) {
                            const T a0 = r[ia], a1 = r[6 + ia], a2 = r[12 + ia];
#pragma unroll
                            for (int p = 0; p < 3; ++p) {
                                if (p == 2 && hl >= 4) break;
                                const int yr = ry + pib[p], zr = rz + pic[p];
                                if ((unsigned)yr < (unsigned)bh && (unsigned)zr < (unsigned)bd) {
                                    const T* s6 = r + 18 + pib[p] * 6;
                                    const T t0 = a0 * s6[0] + a1 * s6[1] + a2 * s6[2];
                                    const T t1 = a0 * s6[3] + a1 * s6[4];
                                    const T t2 = a0 * s6[5];
                                    tplane[yr * BZP + zr] += t0 * r[54 + pic[p]] + t1 * r[60 + pic[p]] + t2 * r[66 + pic[p]];
                                }
                            }
                        } else {
                            const T a0 = r[ia];
#pragma unroll
                            for (int p = 0; p < 3; ++p) {
                                if (p == 2 && hl >= 4) break;
                                const int yr = ry + pib[p], zr = rz + pic[p];
                                if ((unsigned)yr < (unsigned)bh && (unsigned)zr < (unsigned)bd)
                                    tplane[yr * BZP + zr] += a0 * r[6 + pib[p]] * r[12 + pic[p]];
                            }
                        }
                    }
                    __syncwarp();                        // lanes of a half-warp hand tile addresses to each other between atoms
                }
            }
            __syncthreads();
        }
    }
    // ---- write-out: every mesh point of the brick exactly once, full rows
    for (int r0 = warp * RPW; r0 < BRX * BRY; r0 += BR_WARPS * RPW) {
        const int r = r0 + lane / BZ, x = r / BRY, y = r - x * BRY;
        if (x < bw && y < bh && zl < bd) mesh[((size_t)(x0 + x) * K2 + (y0 + y)) * K3 + z0 + zl] = tile[(x * BRY + y) * BZP + zl];
    }
}

// ---------------------------------------------------------------------------- host side
bool brick_supported(const int K[3]) { return K[0] >= 3 * BRX && K[1] >= 3 * BRY && K[2] >= 3 * 16; }

template <typename T, bool MP, int BZ>
static cudaError_t brick_attr() {
    return cudaFuncSetAttribute(spread_brick_kernel<T, MP, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)BrickSmem<T, MP, BZ>::bytes);
}

cudaError_t brick_alloc(BrickWork& w, int n_atoms, const int K[3], int n_sm, size_t elem_bytes) {
    brick_free(w);
    if (!brick_supported(K) || n_atoms <= 0) return cudaSuccess;
    int bz = 16;
    const long long nb32 = (long long)((K[0] + BRX - 1) / BRX) * ((K[1] + BRY - 1) / BRY) * ((K[2] + 31) / 32);
    if (K[2] >= 96 && nb32 >= 8LL * n_sm) bz = 32;       // large meshes: fewer, longer rows (256-byte stores, 1.99 vs 2.26 visits per atom)
    if (const char* e = getenv("ADMP_BRICK_Z")) {
        const int v = atoi(e);
        if (v == 16 || (v == 32 && K[2] >= 96)) bz = v;
    }
    w.geom.bz = bz;
    w.geom.nb[0] = (K[0] + BRX - 1) / BRX;
    w.geom.nb[1] = (K[1] + BRY - 1) / BRY;
    w.geom.nb[2] = (K[2] + bz - 1) / bz;
    w.n_bricks = w.geom.nb[0] * w.geom.nb[1] * w.geom.nb[2];
    w.n_atoms = n_atoms;
    w.rec_bytes = (size_t)n_atoms * BrickRec<true>::RS * elem_bytes;
    cudaError_t e;
    if ((e = cudaMalloc(&w.count, sizeof(int32_t) * ((size_t)w.n_bricks + 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.start, sizeof(int32_t) * ((size_t)w.n_bricks + 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.tmp, sizeof(int4) * (size_t)n_atoms)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.anchor, sizeof(int4) * (size_t)n_atoms)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.rec, w.rec_bytes)) != cudaSuccess) return e;
    if ((e = brick_attr<double, true, 16>()) != cudaSuccess) return e;
    if ((e = brick_attr<double, false, 16>()) != cudaSuccess) return e;
    if ((e = brick_attr<double, true, 32>()) != cudaSuccess) return e;
    if ((e = brick_attr<double, false, 32>()) != cudaSuccess) return e;
    if ((e = brick_attr<float, true, 16>()) != cudaSuccess) return e;
    if ((e = brick_attr<float, false, 16>()) != cudaSuccess) return e;
    if ((e = brick_attr<float, true, 32>()) != cudaSuccess) return e;
    if ((e = brick_attr<float, false, 32>()) != cudaSuccess) return e;
    w.ready = true;
    return cudaSuccess;
}

void brick_free(BrickWork& w) {
    if (w.count) cudaFree(w.count);
    if (w.start) cudaFree(w.start);
    if (w.tmp) cudaFree(w.tmp);
    if (w.anchor) cudaFree(w.anchor);
    if (w.rec) cudaFree(w.rec);
    w = BrickWork{};
}

size_t brick_bytes(const BrickWork& w) {
    return w.ready ? 2 * sizeof(int32_t) * ((size_t)w.n_bricks + 1) + 2 * sizeof(int4) * (size_t)w.n_atoms + w.rec_bytes : 0;
}

template <typename T>
void launch_brick_sort(cudaStream_t st, const BrickWork& w, const BoxInfo* B, const void* pos) {
    const int n = w.n_atoms;
    cudaMemsetAsync(w.count, 0, sizeof(int32_t) * ((size_t)w.n_bricks + 1), st);
    brick_count_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, B, (const T*)pos, w.geom, w.tmp, w.count);
    brick_scan_kernel<<<1, 1024, 0, st>>>(w.n_bricks, w.count, w.start);
    brick_fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, w.tmp, w.start, w.count, w.anchor);
}

template <typename T>
void launch_spread_brick(cudaStream_t st, const BrickWork& w, const BoxInfo* B, const void* pos, const void* M, int m_cols,
                         int m_stride, const void* U, void* mesh) {
    const int n = w.n_atoms;
    const unsigned pgrid = (n + PREP_ATOMS - 1) / PREP_ATOMS;
#define ADMP_B_LAUNCH(MP, BZ, u)                                                                                                    \
    do {                                                                                                                            \
        brick_prep_kernel<T, MP><<<pgrid, PREP_THREADS, 0, st>>>(n, B, w.anchor, (const T*)pos, (const T*)M, m_stride, u, (T*)w.rec); \
        spread_brick_kernel<T, MP, BZ><<<w.n_bricks, BR_THREADS, BrickSmem<T, MP, BZ>::bytes, st>>>(B, w.geom, w.start, w.anchor,   \
                                                                                                    (const T*)w.rec, (T*)mesh);     \
    } while (0)
    if (w.geom.bz == 32) {
        if (m_cols >= 10) ADMP_B_LAUNCH(true, 32, (const T*)U); else ADMP_B_LAUNCH(false, 32, nullptr);
    } else {
        if (m_cols >= 10) ADMP_B_LAUNCH(true, 16, (const T*)U); else ADMP_B_LAUNCH(false, 16, nullptr);
    }
#undef ADMP_B_LAUNCH
}

template void launch_brick_sort<double>(cudaStream_t, const BrickWork&, const BoxInfo*, const void*);
template void launch_brick_sort<float>(cudaStream_t, const BrickWork&, const BoxInfo*, const void*);
template void launch_spread_brick<double>(cudaStream_t, const BrickWork&, const BoxInfo*, const void*, const void*, int, int,
                                          const void*, void*);
template void launch_spread_brick<float>(cudaStream_t, const BrickWork&, const BoxInfo*, const void*, const void*, int, int,
                                         const void*, void*);

}  // namespace admp
