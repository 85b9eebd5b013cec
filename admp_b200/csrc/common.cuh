// Shared device helpers for libadmp_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/admp_b200.h"

#define ADMP_DIEL 1389.35455846            // admp/pme.py:16
#define ADMP_THOLE_DEFAULT 0.3             // admp/pme.py:17
#define ADMP_SQRT_PI 1.7724538509055159    // admp/recip.py:19
#define ADMP_SQRT3 1.7320508075688772

namespace admp {

// Everything the kernels need to know about the cell, derived on the device from the
// caller's box (so a changing box never forces a host round trip). Always double.
struct BoxInfo {
    double box[9];    // rows = lattice vectors
    double inv[9];    // inverse
    double nstar[9];  // nstar[d][c] = K_d * inv[c][d]            (admp/recip.py:55)
    double vol;       // det(box)
    int K[3];
};

// x-slab decomposition of the mesh / half spectrum over the GPUs of one NVLink domain: rank r owns the planes
// [start[r], start[r+1]) (start[r] = r*K1/n, so any K1 works); plane i1 of the (K1, K2, K3) array lives at the usual
// offset ((i1*K2 + i2)*K3 + i3) inside base[owner(i1)], where base[r] is rank r's buffer mapped into this process
// (cudaIpc peer mapping, or a plain pointer for r == own rank).
constexpr int ADMP_MAX_PEERS = 8;
struct PeerTab {
    void* base[ADMP_MAX_PEERS];
    int start[ADMP_MAX_PEERS + 1];
    int n;         // ranks
    __host__ __device__ __forceinline__ int owner(int plane) const {
        int o = 0;
#pragma unroll
        for (int r = 1; r < ADMP_MAX_PEERS; ++r) o += (r < n && plane >= start[r]) ? 1 : 0;
        return o;
    }
    __host__ __device__ __forceinline__ int planes(int r) const { return start[r + 1] - start[r]; }
};

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NV doubles held per thread; thread 0 adds the totals to dst[0..NV).
// smem must hold NV * (blockDim.x/32) doubles. All threads of the block must call.
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double* smem, double* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) smem[k * nwarp + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarp; ++w) s += smem[threadIdx.x * nwarp + w];
        if (s != 0.0) atomicAdd(dst + threadIdx.x, s);
    }
    __syncthreads();
}

// atomic max for non-negative doubles through their (order-preserving) bit pattern
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}

// minimum image of d (Cartesian) : ds = d.inv ; ds -= floor(ds + 0.5) ; d = ds.box
// (admp/spatial.py:29-32). The integer image vector is returned for the box adjoint.
template <typename T>
__device__ __forceinline__ void min_image(const BoxInfo& B, T (&d)[3], T (&sh)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T s = d[0] * (T)B.inv[0 * 3 + a] + d[1] * (T)B.inv[1 * 3 + a] + d[2] * (T)B.inv[2 * 3 + a];
        sh[a] = floor(s + (T)0.5);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        d[c] -= sh[0] * (T)B.box[0 * 3 + c] + sh[1] * (T)B.box[1 * 3 + c] + sh[2] * (T)B.box[2 * 3 + c];
}

// bond count between i and j from the CSR covalent map, mapped to the reference's scale
// index: Scales[covalent_map[i,j]-1] with 0 -> -1 -> last entry (admp/pme.py:681-683, A2).
__device__ __forceinline__ int scale_index(const int32_t* __restrict__ off, const int32_t* __restrict__ idx,
                                           const int8_t* __restrict__ nb, int i, int j) {
    int n = 0;
    if (off != nullptr) {
        for (int k = off[i], e = off[i + 1]; k < e; ++k)
            if (idx[k] == j) { n = nb[k]; break; }
    }
    int s = n - 1;
    if (s < 0) s = 4;
    if (s > 4) s = 4;     // jnp clamps out-of-range gathers
    return s;
}

}  // namespace admp
