// Generic pair driver (admp/pairwise.py:45-91): the geometry half of generate_pairwise_interaction for ANY per-pair
// energy kernel. The reference gathers positions per pair, applies the minimum image, takes the norm and hands
// (dr, mscale, per-pair parameters) to a user kernel; jax.grad then pushes dE/d(dr) back to the positions and the box.
// Here the two halves that touch per-atom arrays are CUDA kernels:
//   pair_geom_kernel      rows -> dr = |min_image(r_i - r_j)| and the scale index covalent_map[i,j]-1 (0 -> last entry, A2)
//   pair_geom_bwd_kernel  dE/d(dr) per row -> dE/dpositions (+= on i, -= on j) and the image-shift part of dE/dbox
// and the user kernel itself runs between them as an element-wise function of device tensors (admp_b200/pairwise.py).
// Kernels with a fused device body (the reference's TT_damping_qq_c6_kernel, pair.cu) bypass this path.
#include "kernels.h"

namespace admp {

template <typename T>
__global__ void __launch_bounds__(256)
pair_geom_kernel(int64_t n_rows, int n_atoms, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos, const int32_t* __restrict__ pairs,
                 const int32_t* __restrict__ cov_off, const int32_t* __restrict__ cov_idx, const int8_t* __restrict__ cov_nb,
                 T* __restrict__ dr, int32_t* __restrict__ sidx) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rows) return;
    const int i = pairs[2 * p], j = pairs[2 * p + 1];
    const bool live = (i < j) && (i >= 0) && (j < n_atoms);                 // pairwise.py:60 (padding rows are (N, N))
    T r = (T)1;
    int s = -1;
    if (live) {
        T d[3] = {pos[3 * (size_t)i] - pos[3 * (size_t)j], pos[3 * (size_t)i + 1] - pos[3 * (size_t)j + 1], pos[3 * (size_t)i + 2] - pos[3 * (size_t)j + 2]};
        T sh[3];
        min_image(*Bp, d, sh);
        r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        s = scale_index(cov_off, cov_idx, cov_nb, i, j);
    }
    dr[p] = r;
    sidx[p] = s;
}

template <typename T>
__global__ void __launch_bounds__(128)
pair_geom_bwd_kernel(int64_t n_rows, int n_atoms, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos, const int32_t* __restrict__ pairs,
                     const T* __restrict__ gdr, uint32_t flags, T* __restrict__ dpos, double* __restrict__ scalars) {
    __shared__ double red[9 * 4];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (p < n_rows) {
        const int i = pairs[2 * p], j = pairs[2 * p + 1];
        if ((i < j) && (i >= 0) && (j < n_atoms)) {
            T d[3] = {pos[3 * (size_t)i] - pos[3 * (size_t)j], pos[3 * (size_t)i + 1] - pos[3 * (size_t)j + 1], pos[3 * (size_t)i + 2] - pos[3 * (size_t)j + 2]};
            T sh[3];
            min_image(*Bp, d, sh);
            const T rinv = rsqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            const T g = gdr[p] * rinv;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const T f = g * d[k];
                atomicAdd(dpos + (size_t)i * 3 + k, f);
                atomicAdd(dpos + (size_t)j * 3 + k, -f);
                if (flags & ADMP_WANT_VIRIAL) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) acc_box[3 * a + k] = -(double)(sh[a] * f);
                }
            }
        }
    }
    if (flags & ADMP_WANT_VIRIAL) block_accumulate<9>(acc_box, red, scalars + ADMP_S_DBOX);
}

template <typename T>
void launch_pair_geom(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                      const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb, void* dr, int32_t* sidx) {
    if (n_rows <= 0) return;
    pair_geom_kernel<T><<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(n_rows, n_atoms, B, (const T*)pos, pairs, cov_off, cov_idx, cov_nb,
                                                                         (T*)dr, sidx);
}
template <typename T>
void launch_pair_geom_bwd(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                          const void* gdr, uint32_t flags, void* dpos, double* scalars) {
    if (n_rows <= 0) return;
    pair_geom_bwd_kernel<T><<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(n_rows, n_atoms, B, (const T*)pos, pairs, (const T*)gdr, flags,
                                                                             (T*)dpos, scalars);
}
#define ADMP_INST(T)                                                                                                                   \
    template void launch_pair_geom<T>(cudaStream_t, int64_t, int, const BoxInfo*, const void*, const int32_t*, const int32_t*,           \
                                      const int32_t*, const int8_t*, void*, int32_t*);                                                  \
    template void launch_pair_geom_bwd<T>(cudaStream_t, int64_t, int, const BoxInfo*, const void*, const int32_t*, const void*, uint32_t, \
                                          void*, double*);
ADMP_INST(double)
ADMP_INST(float)
#undef ADMP_INST

}  // namespace admp
