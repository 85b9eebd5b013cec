// GPU cell-list neighbour build, the replacement for the jax_md call every reference script
// makes (examples/water_1024/run_admp.py:109-112; jax_md is third-party, version unpinned):
//   pairs = {(i,j) : i<j, |min_image(r_i - r_j)|^2 < rc^2}, rows sorted by (i,j), padded (N,N).
//
// The keep/drop predicate is evaluated in float64 with a FIXED operation order and no fused
// multiply-add, identical to oracle/pairlist.py, so that the pair SET is bit-exact:
//   s = r*(1/L) ; t = (s_i - s_j) + 0.5 ; ds = (t - floor(t)) - 0.5 ; d = ds*L
//   d2 = (dx*dx + dy*dy) + dz*dz ; keep iff d2 < rc*rc
// Orthorhombic boxes (the scope of the reference's examples). HBM-bound: Na*3*w read, Np*8 written.
#include <cstdlib>
#include <cstring>

#include <cub/device/device_scan.cuh>

#include "kernels.h"

namespace admp {

struct NbGeom {
    double L[3], invL[3], rc2;
    int nc[3];
};

__device__ __forceinline__ double frac_coord(double r, double invL) { return __dmul_rn(r, invL); }

__device__ __forceinline__ bool pair_within(const NbGeom& g, const double (&si)[3], const double (&sj)[3]) {
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double t = __dadd_rn(__dsub_rn(si[k], sj[k]), 0.5);
        const double ds = __dsub_rn(__dsub_rn(t, floor(t)), 0.5);
        d[k] = __dmul_rn(ds, g.L[k]);
    }
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
    return d2 < g.rc2;
}

template <typename T>
__global__ void nb_geom_kernel(const BoxInfo* __restrict__ B, double rc, int ncx, int ncy, int ncz, NbGeom* __restrict__ g) {
    if (threadIdx.x || blockIdx.x) return;
    for (int k = 0; k < 3; ++k) { g->L[k] = B->box[4 * k]; g->invL[k] = 1.0 / B->box[4 * k]; }
    g->rc2 = __dmul_rn(rc, rc);
    g->nc[0] = ncx; g->nc[1] = ncy; g->nc[2] = ncz;
}

template <typename T>
__device__ __forceinline__ void load_s(const NbGeom& g, const T* __restrict__ pos, int a, double (&s)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = frac_coord((double)pos[3 * a + k], g.invL[k]);
}

__device__ __forceinline__ int cell_coord(double s, int nc) {
    double w = s - floor(s);
    int c = (int)(w * nc);
    return c >= nc ? nc - 1 : (c < 0 ? 0 : c);
}

template <typename T>
__global__ void nb_assign_kernel(int n, const NbGeom* __restrict__ gp, const T* __restrict__ pos, int32_t* __restrict__ cell_of,
                                 int32_t* __restrict__ cell_count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const NbGeom& g = *gp;
    double s[3];
    load_s(g, pos, a, s);
    const int c = (cell_coord(s[0], g.nc[0]) * g.nc[1] + cell_coord(s[1], g.nc[1])) * g.nc[2] + cell_coord(s[2], g.nc[2]);
    cell_of[a] = c;
    atomicAdd(cell_count + c, 1);
}

__global__ void nb_fill_kernel(int n, const int32_t* __restrict__ cell_of, const int32_t* __restrict__ cell_start,
                               int32_t* __restrict__ cursor, int32_t* __restrict__ sorted) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int c = cell_of[a];
    sorted[cell_start[c] + atomicAdd(cursor + c, 1)] = a;
}

// deterministic order inside each cell (ascending atom index)
__global__ void nb_sort_cells_kernel(int ncell, const int32_t* __restrict__ cell_start, int32_t* __restrict__ sorted) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    for (int i = b + 1; i < e; ++i) {
        const int v = sorted[i];
        int j = i - 1;
        while (j >= b && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; }
        sorted[j + 1] = v;
    }
}

// EMIT = false: count neighbours j>i of atom i. EMIT = true: write them (ascending j) at nbr_start[i].
template <typename T, bool EMIT>
__global__ void __launch_bounds__(128)
nb_pairs_kernel(int n, const NbGeom* __restrict__ gp, const T* __restrict__ pos, const int32_t* __restrict__ cell_start,
                const int32_t* __restrict__ sorted, int32_t* __restrict__ nbr_count, const int32_t* __restrict__ nbr_start,
                int32_t* __restrict__ pairs, int64_t capacity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const NbGeom& g = *gp;
    double si[3];
    load_s(g, pos, i, si);
    int ci[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ci[k] = cell_coord(si[k], g.nc[k]);
    int cnt = 0;
    const int64_t base = EMIT ? (int64_t)nbr_start[i] : 0;
    // neighbour cells, each visited once even when a dimension has fewer than 3 cells
    for (int ox = 0; ox < (g.nc[0] < 3 ? g.nc[0] : 3); ++ox)
        for (int oy = 0; oy < (g.nc[1] < 3 ? g.nc[1] : 3); ++oy)
            for (int oz = 0; oz < (g.nc[2] < 3 ? g.nc[2] : 3); ++oz) {
                const int cx = g.nc[0] < 3 ? ox : (ci[0] + ox - 1 + g.nc[0]) % g.nc[0];
                const int cy = g.nc[1] < 3 ? oy : (ci[1] + oy - 1 + g.nc[1]) % g.nc[1];
                const int cz = g.nc[2] < 3 ? oz : (ci[2] + oz - 1 + g.nc[2]) % g.nc[2];
                const int c = (cx * g.nc[1] + cy) * g.nc[2] + cz;
                for (int k = cell_start[c], e = cell_start[c + 1]; k < e; ++k) {
                    const int j = sorted[k];
                    if (j <= i) continue;
                    double sj[3];
                    load_s(g, pos, j, sj);
                    if (pair_within(g, si, sj)) {
                        if (EMIT) {
                            const int64_t row = base + cnt;
                            if (row < capacity) { pairs[2 * row] = i; pairs[2 * row + 1] = j; }
                        }
                        ++cnt;
                    }
                }
            }
    if (!EMIT) { nbr_count[i] = cnt; return; }
    // ascending j within the rows of atom i (rows beyond capacity were not written)
    int64_t m = cnt;
    if (base + m > capacity) m = capacity > base ? capacity - base : 0;
    for (int64_t a = 1; a < m; ++a) {
        const int v = pairs[2 * (base + a) + 1];
        int64_t b = a - 1;
        while (b >= 0 && pairs[2 * (base + b) + 1] > v) { pairs[2 * (base + b + 1) + 1] = pairs[2 * (base + b) + 1]; --b; }
        pairs[2 * (base + b + 1) + 1] = v;
    }
}

// The same two passes with ONE WARP PER ATOM (the per-thread walk above is a serial chain of ~27 cells x dependent loads: 85 us
// per pass for 3 072 atoms): lanes = the <= 27 neighbour cells (start / count, warp scan), then the candidates 32 at a time
// (cell by binary search in the prefix, float64 predicate per lane, ballot -> running count). EMIT writes the hits unsorted
// into column 0 of the atom's rows, ranks every hit among the row segment (ascending j, the order jax_md's OrderedSparse emits
// and the cluster pair kernel needs) and writes (i, j) - the pair SET and the row order are exactly those of nb_pairs_kernel.
constexpr int NBW_WARPS = 4;
template <typename T, bool EMIT>
__global__ void __launch_bounds__(NBW_WARPS * 32)
nb_pairs_warp_kernel(int n, const NbGeom* __restrict__ gp, const T* __restrict__ pos, const int32_t* __restrict__ cell_start,
                     const int32_t* __restrict__ sorted, int32_t* __restrict__ nbr_count, const int32_t* __restrict__ nbr_start,
                     int32_t* pairs, int64_t capacity) {
    __shared__ int s_lo[NBW_WARPS][32], s_pref[NBW_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * NBW_WARPS + warp;
    if (i >= n) return;                                   // whole warps leave together
    const NbGeom& g = *gp;
    double si[3];
    load_s(g, pos, i, si);
    int ci[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ci[k] = cell_coord(si[k], g.nc[k]);
    const int nx = g.nc[0] < 3 ? g.nc[0] : 3, ny = g.nc[1] < 3 ? g.nc[1] : 3, nz = g.nc[2] < 3 ? g.nc[2] : 3;
    const int ncell = nx * ny * nz;
    int lo = 0, cnt = 0;
    if (lane < ncell) {
        const int oz = lane % nz, oy = (lane / nz) % ny, ox = lane / (nz * ny);
        const int cx = g.nc[0] < 3 ? ox : (ci[0] + ox - 1 + g.nc[0]) % g.nc[0];
        const int cy = g.nc[1] < 3 ? oy : (ci[1] + oy - 1 + g.nc[1]) % g.nc[1];
        const int cz = g.nc[2] < 3 ? oz : (ci[2] + oz - 1 + g.nc[2]) % g.nc[2];
        const int c = (cx * g.nc[1] + cy) * g.nc[2] + cz;
        lo = cell_start[c];
        cnt = cell_start[c + 1] - lo;
    }
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += y;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    s_lo[warp][lane] = lo;
    s_pref[warp][lane] = pre - cnt;                       // exclusive prefix (lanes >= ncell: total)
    __syncwarp();
    const int64_t base = EMIT ? (int64_t)nbr_start[i] : 0;
    const unsigned lt = (1u << lane) - 1u;
    int m = 0;
    for (int q0 = 0; q0 < total; q0 += 32) {
        const int q = q0 + lane;
        bool ok = false;
        int j = -1;
        if (q < total) {
            int c = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (c + step < ncell && s_pref[warp][c + step] <= q) c += step;
            j = sorted[s_lo[warp][c] + (q - s_pref[warp][c])];
            if (j > i) {
                double sj[3];
                load_s(g, pos, j, sj);
                ok = pair_within(g, si, sj);
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (EMIT && ok) {
            const int64_t row = base + m + __popc(mask & lt);
            if (row < capacity) pairs[2 * row] = j;       // column 0 = scratch until the ranks are known
        }
        m += __popc(mask);
    }
    if (!EMIT) {
        if (lane == 0) nbr_count[i] = m;
        return;
    }
    int64_t mm = m;
    if (base + mm > capacity) mm = capacity > base ? capacity - base : 0;
    __syncwarp();
    // rank of every hit among the row segment (all j distinct) -> ascending j in column 1; then column 0 = i
    for (int64_t a0 = 0; a0 < mm; a0 += 32) {
        const int64_t a = a0 + lane;
        if (a < mm) {
            const int v = __ldcg(pairs + 2 * (base + a));
            int rank = 0;
            for (int64_t b = 0; b < mm; ++b) rank += (__ldcg(pairs + 2 * (base + b)) < v) ? 1 : 0;
            pairs[2 * (base + rank) + 1] = v;
        }
    }
    __syncwarp();
    for (int64_t a = lane; a < mm; a += 32) pairs[2 * (base + a)] = i;
}

__global__ void nb_pad_kernel(int n, const int32_t* __restrict__ nbr_start, int32_t* __restrict__ pairs, int64_t capacity,
                              int32_t* __restrict__ info) {
    const int64_t total = nbr_start[n];
    const int64_t first = total < capacity ? total : capacity;
    for (int64_t r = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < capacity; r += (int64_t)gridDim.x * blockDim.x) {
        pairs[2 * r] = n; pairs[2 * r + 1] = n;                     // jax_md pads with (N, N)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = (int32_t)total; info[1] = total > capacity ? 1 : 0; }
}

// geometry block + CUB scratch live in the context's NbWork (one per admp_ctx: no sharing across contexts, streams or devices)
static cudaError_t exclusive_scan(cudaStream_t st, NbWork& w, const int32_t* in, int32_t* out, int n) {
    size_t need = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n, st);
    if (e != cudaSuccess) return e;
    if (need > w.scan_tmp_bytes) {
        if (w.scan_tmp) cudaFree(w.scan_tmp);
        w.scan_tmp = nullptr;
        w.scan_tmp_bytes = 0;
        e = cudaMalloc(&w.scan_tmp, need);
        if (e != cudaSuccess) return e;
        w.scan_tmp_bytes = need;
    }
    return cub::DeviceScan::ExclusiveSum(w.scan_tmp, need, in, out, n, st);
}

cudaError_t launch_nblist(cudaStream_t st, const BoxInfo* B, const void* pos, int dtype, int n, double rc, NbWork& w, int ncx, int ncy,
                          int ncz, int32_t* pairs, int64_t capacity, int32_t* info) {
    if (w.geom == nullptr) {
        cudaError_t e = cudaMalloc(&w.geom, sizeof(NbGeom));
        if (e != cudaSuccess) return e;
    }
    NbGeom* g_geom = static_cast<NbGeom*>(w.geom);
    cudaError_t e;
    const int ncell = ncx * ncy * ncz;
    const int tb = 128, gb = (n + tb - 1) / tb;
    cudaMemsetAsync(w.cell_count, 0, sizeof(int32_t) * (ncell + 1), st);
    nb_geom_kernel<double><<<1, 32, 0, st>>>(B, rc, ncx, ncy, ncz, g_geom);
    if (dtype == ADMP_F64) nb_assign_kernel<double><<<gb, tb, 0, st>>>(n, g_geom, (const double*)pos, w.cell_of, w.cell_count);
    else nb_assign_kernel<float><<<gb, tb, 0, st>>>(n, g_geom, (const float*)pos, w.cell_of, w.cell_count);
    if ((e = exclusive_scan(st, w, w.cell_count, w.cell_start, ncell + 1)) != cudaSuccess) return e;
    cudaMemsetAsync(w.cell_count, 0, sizeof(int32_t) * (ncell + 1), st);       // reused as the fill cursor
    nb_fill_kernel<<<gb, tb, 0, st>>>(n, w.cell_of, w.cell_start, w.cell_count, w.sorted);
    // one warp per atom up to 2^18 atoms (3 072 atoms: 209 -> 92 us per list; 98 304 atoms at liquid density, rc 8 A: 3.9 -> 1.5 ms);
    // beyond that one thread per atom already fills the GPU and wins (786 432 gas-like atoms: 0.80 vs 1.11 ms). ADMP_NBLIST=thread|warp
    static const int forced = [] { const char* e = getenv("ADMP_NBLIST"); return !e ? 0 : (strcmp(e, "thread") == 0 ? 1 : (strcmp(e, "warp") == 0 ? 2 : 0)); }();
    const bool per_thread = forced == 1 || (forced == 0 && n > (1 << 18));
    cudaMemsetAsync(w.nbr_count, 0, sizeof(int32_t) * (n + 1), st);
    if (per_thread) {
        nb_sort_cells_kernel<<<(ncell + tb - 1) / tb, tb, 0, st>>>(ncell, w.cell_start, w.sorted);
        if (dtype == ADMP_F64)
            nb_pairs_kernel<double, false><<<gb, tb, 0, st>>>(n, g_geom, (const double*)pos, w.cell_start, w.sorted, w.nbr_count, nullptr, nullptr, 0);
        else
            nb_pairs_kernel<float, false><<<gb, tb, 0, st>>>(n, g_geom, (const float*)pos, w.cell_start, w.sorted, w.nbr_count, nullptr, nullptr, 0);
        if ((e = exclusive_scan(st, w, w.nbr_count, w.nbr_start, n + 1)) != cudaSuccess) return e;
        if (dtype == ADMP_F64)
            nb_pairs_kernel<double, true><<<gb, tb, 0, st>>>(n, g_geom, (const double*)pos, w.cell_start, w.sorted, nullptr, w.nbr_start, pairs, capacity);
        else
            nb_pairs_kernel<float, true><<<gb, tb, 0, st>>>(n, g_geom, (const float*)pos, w.cell_start, w.sorted, nullptr, w.nbr_start, pairs, capacity);
    } else {
        const int gw = (n + NBW_WARPS - 1) / NBW_WARPS, tw = NBW_WARPS * 32;
        if (dtype == ADMP_F64)
            nb_pairs_warp_kernel<double, false><<<gw, tw, 0, st>>>(n, g_geom, (const double*)pos, w.cell_start, w.sorted, w.nbr_count, nullptr, nullptr, 0);
        else
            nb_pairs_warp_kernel<float, false><<<gw, tw, 0, st>>>(n, g_geom, (const float*)pos, w.cell_start, w.sorted, w.nbr_count, nullptr, nullptr, 0);
        if ((e = exclusive_scan(st, w, w.nbr_count, w.nbr_start, n + 1)) != cudaSuccess) return e;
        if (dtype == ADMP_F64)
            nb_pairs_warp_kernel<double, true><<<gw, tw, 0, st>>>(n, g_geom, (const double*)pos, w.cell_start, w.sorted, nullptr, w.nbr_start, pairs, capacity);
        else
            nb_pairs_warp_kernel<float, true><<<gw, tw, 0, st>>>(n, g_geom, (const float*)pos, w.cell_start, w.sorted, nullptr, w.nbr_start, pairs, capacity);
    }
    nb_pad_kernel<<<64, 256, 0, st>>>(n, w.nbr_start, pairs, capacity, info);
    return cudaGetLastError();
}

}  // namespace admp
