// Micro-benchmark: FP64 / FP32 FMA throughput and dependent-issue latency on the device
// (the FP-pipe roofline denominators for the pair kernel and the FFT butterflies).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fp_peak tools/fp_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
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
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T, int ILP>
double run(int blocks, int threads, int iters, int n_sm) {
    T* out;
    cudaMalloc(&out, sizeof(T) * blocks * threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    return 2.0 * ILP * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s: %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, clk / 1e3);
    const int sm = p.multiProcessorCount;
    printf("FP64 FMA TFLOP/s  (blocks/SM x threads, ILP)\n");
    printf("  full  8x256 ILP8 : %.2f\n", run<double, 8>(sm * 8, 256, 4096, sm));
    printf("  full  4x256 ILP4 : %.2f\n", run<double, 4>(sm * 4, 256, 4096, sm));
    printf("  1 warp/SMSP (1x128) ILP1  : %.3f\n", run<double, 1>(sm, 128, 16384, sm));
    printf("  1 warp/SMSP (1x128) ILP2  : %.3f\n", run<double, 2>(sm, 128, 16384, sm));
    printf("  1 warp/SMSP (1x128) ILP4  : %.3f\n", run<double, 4>(sm, 128, 16384, sm));
    printf("  1 warp/SMSP (1x128) ILP8  : %.3f\n", run<double, 8>(sm, 128, 16384, sm));
    printf("  1 warp/SMSP (1x128) ILP16 : %.3f\n", run<double, 16>(sm, 128, 8192, sm));
    printf("  3 warps/SMSP (3x128) ILP4 : %.3f\n", run<double, 4>(sm * 3, 128, 16384, sm));
    printf("  3 warps/SMSP (3x128) ILP8 : %.3f\n", run<double, 8>(sm * 3, 128, 16384, sm));
    printf("FP32 FMA TFLOP/s\n");
    printf("  full  8x256 ILP8 : %.2f\n", run<float, 8>(sm * 8, 256, 8192, sm));
    printf("  1 warp/SMSP ILP1 : %.3f\n", run<float, 1>(sm, 128, 32768, sm));
    printf("  1 warp/SMSP ILP8 : %.3f\n", run<float, 8>(sm, 128, 32768, sm));
    return 0;
}
