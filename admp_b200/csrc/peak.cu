// FP-pipe roofline denominator: dense FMA throughput of the FP64 / FP32 CUDA-core pipes, measured on the
// device the library runs on (the pair kernel and the FFT butterflies are bounded by these pipes, not by
// tensor cores). Exposed as admp_fp_peak so bench.py reports `peak` from a live measurement.
#include <cuda_runtime.h>

#include "../../include/admp_b200.h"

namespace {
template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* out, int iters, T a, T b) {
    T v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = fma(v[k], a, b);
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
int run_peak(cudaStream_t st, double* tflops) {
    int dev = 0, nsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = nsm * 8, threads = 256, iters = sizeof(T) == 8 ? 4096 : 16384;
    T* out = nullptr;
    if (cudaMalloc(&out, sizeof(T) * (size_t)blocks * threads) != cudaSuccess) return 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {          // first launch is the warm-up
        cudaEventRecord(a, st);
        fma_chain_kernel<T, 8><<<blocks, threads, 0, st>>>(out, iters, (T)1.0000001, (T)1e-9);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return 1;
    *tflops = 2.0 * 8 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}
}  // namespace

extern "C" int admp_fp_peak(void* stream, int dtype, double* tflops) {
    if (tflops == nullptr) return 1;
    return dtype == ADMP_F32 ? run_peak<float>((cudaStream_t)stream, tflops) : run_peak<double>((cudaStream_t)stream, tflops);
}
