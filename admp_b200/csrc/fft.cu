// Hand-written 3-D real FFT for PME meshes whose sizes cuFFT handles badly, fused with the
// influence-function convolution.
//
// Why: the reference's Ewald set-up gives K = 154 per 50 A (154 = 2*7*11); cuFFT splits each
// transform into 6-7 generic kernels (regular_fft_factor<2|7|11>, pre/postprocess) and the
// separate convolution pass re-reads and re-writes the spectrum. Here one reciprocal round trip is
// five passes over the mesh instead of cuFFT's ~14 kernel launches + 1:
//     Z-forward (R2C)  ->  Y-forward  ->  [X-forward * C_k/theta^2 (+energy, +virial) * X-inverse]
//     ->  Y-inverse  ->  Z-inverse (C2R)
// i.e. 10*w*G bytes of traffic instead of 14*w*G, and the spectrum never exists in its x-transformed
// form outside shared memory.
//
// Each block transforms TL lines in shared memory with a Stockham autosort mixed-radix FFT
// (radices 2,3,4,5,7,11,13; odd radices use the symmetric cos/sin formulation, tables folded to
// immediates). Lines of the Y/X passes are strided in memory: a block takes TL neighbouring lines
// so every global access is a TL*sizeof(complex) contiguous segment. Line stride in shared memory is
// odd (in complex units) to keep the transposing loads bank-conflict free.
// Index math validated by tools/fft_model.py (tests/test_fft_model.py).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>                 // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cudaTypedefs.h>

#include "fft_trig.h"
#include "influence.cuh"
#include "kernels.h"

namespace admp {

template <typename T> struct alignas(2 * sizeof(T)) cx { T x, y; };   // 16-byte (8-byte for float) vector loads / stores
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// Butterfly constants live in constant memory: a float64 literal whose low word is not zero cannot be an immediate of DFMA, so
// as a literal every use costs two UMOV (9 % of the instructions of the 308-point X pass, ncu r2i); as a constant-bank operand
// (DFMA R, R, c[3][..], R) it costs nothing. One table per radix and precision, entry m = cos / sin(2 pi m / R).
#define ADMP_TRIG_TABLE(R)                                                                                                        \
    __constant__ double g_cos##R##_d[R] = {}; __constant__ double g_sin##R##_d[R] = {};                                           \
    __constant__ float g_cos##R##_f[R] = {};  __constant__ float g_sin##R##_f[R] = {};
ADMP_TRIG_TABLE(3) ADMP_TRIG_TABLE(5) ADMP_TRIG_TABLE(7) ADMP_TRIG_TABLE(11) ADMP_TRIG_TABLE(13) ADMP_TRIG_TABLE(16)
#undef ADMP_TRIG_TABLE
template <typename T, int R> __device__ __forceinline__ T ktrig_cos(int m) {
    if constexpr (sizeof(T) == 8) {
        if constexpr (R == 3) return g_cos3_d[m]; else if constexpr (R == 5) return g_cos5_d[m]; else if constexpr (R == 7) return g_cos7_d[m];
        else if constexpr (R == 11) return g_cos11_d[m]; else if constexpr (R == 13) return g_cos13_d[m]; else return g_cos16_d[m];
    } else {
        if constexpr (R == 3) return g_cos3_f[m]; else if constexpr (R == 5) return g_cos5_f[m]; else if constexpr (R == 7) return g_cos7_f[m];
        else if constexpr (R == 11) return g_cos11_f[m]; else if constexpr (R == 13) return g_cos13_f[m]; else return g_cos16_f[m];
    }
}
template <typename T, int R> __device__ __forceinline__ T ktrig_sin(int m) {
    if constexpr (sizeof(T) == 8) {
        if constexpr (R == 3) return g_sin3_d[m]; else if constexpr (R == 5) return g_sin5_d[m]; else if constexpr (R == 7) return g_sin7_d[m];
        else if constexpr (R == 11) return g_sin11_d[m]; else if constexpr (R == 13) return g_sin13_d[m]; else return g_sin16_d[m];
    } else {
        if constexpr (R == 3) return g_sin3_f[m]; else if constexpr (R == 5) return g_sin5_f[m]; else if constexpr (R == 7) return g_sin7_f[m];
        else if constexpr (R == 11) return g_sin11_f[m]; else if constexpr (R == 13) return g_sin13_f[m]; else return g_sin16_f[m];
    }
}
// filled once per process and device by fft3d_create (host-computed in long double)
static cudaError_t upload_trig_tables() {
    cudaError_t e = cudaSuccess;
#define ADMP_UP(R)                                                                                             \
    {                                                                                                          \
        double c[R], s[R]; float cf[R], sf[R];                                                                 \
        for (int m = 0; m < R; ++m) {                                                                          \
            const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)R; \
            c[m] = (double)cosl(a); s[m] = (double)sinl(a); cf[m] = (float)c[m]; sf[m] = (float)s[m];          \
        }                                                                                                      \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_cos##R##_d, c, sizeof(c));                              \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sin##R##_d, s, sizeof(s));                              \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_cos##R##_f, cf, sizeof(cf));                            \
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sin##R##_f, sf, sizeof(sf));                            \
    }
    ADMP_UP(3) ADMP_UP(5) ADMP_UP(7) ADMP_UP(11) ADMP_UP(13) ADMP_UP(16)
#undef ADMP_UP
    return e;
}

// r-point DFT, SIGN = +1: e^{-i..} (forward), -1: e^{+i..} (inverse)
template <typename T, int R, int SIGN> struct Dft;

template <typename T, int SIGN> struct Dft<T, 2, SIGN> {
    static __device__ __forceinline__ void run(cx<T> (&v)[2]) {
        const cx<T> a = v[0], b = v[1];
        v[0] = a + b; v[1] = a - b;
    }
};
template <typename T, int SIGN> struct Dft<T, 4, SIGN> {
    static __device__ __forceinline__ void run(cx<T> (&v)[4]) {
        const cx<T> s02 = v[0] + v[2], d02 = v[0] - v[2], s13 = v[1] + v[3], d13 = v[1] - v[3];
        // forward: X1 = d02 - i d13, X3 = d02 + i d13
        const cx<T> id13 = {-(T)SIGN * -d13.y, (T)SIGN * -d13.x};   // (-i*SIGN) * d13
        v[0] = s02 + s13; v[2] = s02 - s13;
        v[1] = d02 + id13; v[3] = d02 - id13;
    }
};
// odd radix: X_k, X_{R-k} from a_j = x_j + x_{R-j}, b_j = x_j - x_{R-j}
template <typename T, int R, int SIGN> struct Dft {
    static __device__ __forceinline__ void run(cx<T> (&v)[R]) {
        constexpr int H = (R - 1) / 2;
        cx<T> a[H], b[H];
        cx<T> sum = v[0];
#pragma unroll
        for (int j = 1; j <= H; ++j) { a[j - 1] = v[j] + v[R - j]; b[j - 1] = v[j] - v[R - j]; sum = sum + a[j - 1]; }
        const cx<T> x0 = v[0];
        v[0] = sum;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            T rc = x0.x, ic = x0.y, rs = 0, is = 0;
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const T c = ktrig_cos<T, R>((j * k) % R), s = ktrig_sin<T, R>((j * k) % R);
                rc += a[j - 1].x * c; ic += a[j - 1].y * c;
                rs += b[j - 1].x * s; is += b[j - 1].y * s;
            }
            v[k] = {rc + (T)SIGN * is, ic - (T)SIGN * rs};
            v[R - k] = {rc - (T)SIGN * is, ic + (T)SIGN * rs};
        }
    }
};

template <typename T, int SIGN> struct Dft<T, 8, SIGN> {
    static __device__ __forceinline__ void run(cx<T> (&v)[8]) {
        // two radix-4 on even/odd inputs, then twiddles W8^k and a radix-2 combine
        cx<T> e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
        Dft<T, 4, SIGN>::run(e);
        Dft<T, 4, SIGN>::run(o);
        const T h = ktrig_cos<T, 16>(2);          // sqrt(1/2)
        // W8^1 = (1 - i s)/sqrt2, W8^2 = -i s, W8^3 = (-1 - i s)/sqrt2   with s = SIGN
        const cx<T> o1 = {h * (o[1].x + (T)SIGN * o[1].y), h * (o[1].y - (T)SIGN * o[1].x)};
        const cx<T> o2 = {(T)SIGN * o[2].y, -(T)SIGN * o[2].x};
        const cx<T> o3 = {h * (-o[3].x + (T)SIGN * o[3].y), h * (-o[3].y - (T)SIGN * o[3].x)};
        v[0] = e[0] + o[0]; v[4] = e[0] - o[0];
        v[1] = e[1] + o1;   v[5] = e[1] - o1;
        v[2] = e[2] + o2;   v[6] = e[2] - o2;
        v[3] = e[3] + o3;   v[7] = e[3] - o3;
    }
};
// 16 = 4 x 4 (Cooley-Tukey inside the registers): n = 4 n1 + n2, k = k1 + 4 k2, twiddles W16^(n2 k1)
template <typename T, int SIGN> struct Dft<T, 16, SIGN> {
    static __device__ __forceinline__ cx<T> tw(cx<T> v, int m) {        // v * W16^m, m in {1, 2, 3, 4, 6, 9}
        if (m == 4) return {(T)SIGN * v.y, -(T)SIGN * v.x};            // W16^4 = -i SIGN
        const T wr = ktrig_cos<T, 16>(m), ws = ktrig_sin<T, 16>(m);      // W = cos - i SIGN sin
        if (SIGN > 0) return {v.x * wr + v.y * ws, v.y * wr - v.x * ws};
        return {v.x * wr - v.y * ws, v.y * wr + v.x * ws};
    }
    static __device__ __forceinline__ void run(cx<T> (&v)[16]) {
        cx<T> y[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            cx<T> c[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
            Dft<T, 4, SIGN>::run(c);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = (n2 * k1 == 0) ? c[k1] : tw(c[k1], n2 * k1);
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            cx<T> c[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
            Dft<T, 4, SIGN>::run(c);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = c[k2];
        }
    }
};
// 14 = 2 x 7 by the prime-factor (Good-Thomas) map: no internal twiddles.
//   input  n = (7 n1 + 2 n2) mod 14 ; output k = (7 k1 + 8 k2) mod 14
template <typename T, int SIGN> struct Dft<T, 14, SIGN> {
    static __device__ __forceinline__ void run(cx<T> (&v)[14]) {
        cx<T> a[7], b[7];
#pragma unroll
        for (int n2 = 0; n2 < 7; ++n2) { a[n2] = v[(2 * n2) % 14]; b[n2] = v[(7 + 2 * n2) % 14]; }
        Dft<T, 7, SIGN>::run(a);
        Dft<T, 7, SIGN>::run(b);
#pragma unroll
        for (int k2 = 0; k2 < 7; ++k2) {
            v[(8 * k2) % 14] = a[k2] + b[k2];
            v[(7 + 8 * k2) % 14] = a[k2] - b[k2];
        }
    }
};

// one Stockham stage over TL lines: src -> dst (both [TL][LS] complex in shared memory)
template <typename T, int R, int SIGN>
__device__ __forceinline__ void stage(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst, int N, int TL, int LS, int Ns,
                                      const cx<T>* __restrict__ tw, int twmul) {
    const int m = N / R;
    const int total = TL * m;
    const int step = (m / Ns) * twmul;            // W_N^{t*k*(N/(Ns*R))} = table[t*k*step]
    for (int b = threadIdx.x; b < total; b += blockDim.x) {
        const int l = b / m, j = b - l * m, k = j % Ns;
        const cx<T>* s = src + l * LS + j;
        cx<T> v[R];
        v[0] = s[0];
#pragma unroll
        for (int t = 1; t < R; ++t) {
            cx<T> w = tw[t * k * step];
            if (SIGN < 0) w.y = -w.y;
            v[t] = cmul(s[t * m], w);
        }
        Dft<T, R, SIGN>::run(v);
        cx<T>* d = dst + l * LS + (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) d[t * Ns] = v[t];
    }
}

struct FftPlan {
    int N;          // complex line length
    int nst;
    int radix[12];
};

// in-shared-memory FFT of TL lines; returns the buffer holding the result (A or B)
template <typename T, int SIGN>
__device__ __forceinline__ cx<T>* fft_lines(cx<T>* A, cx<T>* B, const FftPlan& P, int TL, int LS, const cx<T>* __restrict__ tw, int twmul) {
    int Ns = 1;
    for (int s = 0; s < P.nst; ++s) {
        const int r = P.radix[s];
        switch (r) {
            case 2: stage<T, 2, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            case 3: stage<T, 3, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            case 4: stage<T, 4, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            case 5: stage<T, 5, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            case 7: stage<T, 7, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            case 11: stage<T, 11, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
            default: stage<T, 13, SIGN>(A, B, P.N, TL, LS, Ns, tw, twmul); break;
        }
        __syncthreads();
        cx<T>* t = A; A = B; B = t;
        Ns *= r;
    }
    return A;
}

template <typename T>
__device__ __forceinline__ void load_twiddles(cx<T>* stw, const cx<T>* __restrict__ gtw, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) stw[i] = gtw[i];
}

// ------------------------------------------------------------------------------------------ Z passes
// forward: K3 reals per line -> K3/2+1 complex (packed real FFT, tools/fft_model.py r2c)
template <typename T>
__global__ void __launch_bounds__(256)
fft_z_fwd_kernel(FftPlan P, int TL, int LS, int nlines, int K3, const T* __restrict__ mesh, cx<T>* __restrict__ spec,
                 const cx<T>* __restrict__ gtw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* A = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* Bf = A + TL * LS;
    cx<T>* stw = Bf + TL * LS;
    const int M = K3 / 2, K3h = M + 1;
    const int L0 = blockIdx.x * TL;
    const int nl = min(TL, nlines - L0);
    load_twiddles(stw, gtw, K3);
    T* Ar = reinterpret_cast<T*>(A);
    for (int e = threadIdx.x; e < nl * K3; e += blockDim.x) {
        const int l = e / K3, n = e - l * K3;
        Ar[2 * l * LS + n] = mesh[(size_t)(L0 + l) * K3 + n];
    }
    __syncthreads();
    cx<T>* Z = fft_lines<T, 1>(A, Bf, P, nl, LS, stw, 2);       // W_M = W_K3^2
    for (int e = threadIdx.x; e < nl * K3h; e += blockDim.x) {
        const int l = e / K3h, k = e - l * K3h;
        const cx<T> zk = Z[l * LS + (k == M ? 0 : k)];
        cx<T> zc = Z[l * LS + ((k == 0 || k == M) ? 0 : M - k)];
        zc.y = -zc.y;
        const cx<T> a = {(T)0.5 * (zk.x + zc.x), (T)0.5 * (zk.y + zc.y)}, b = {(T)0.5 * (zk.x - zc.x), (T)0.5 * (zk.y - zc.y)};
        const cx<T> w = stw[k];                                   // (cos phi, -sin phi), phi = 2 pi k / K3
        // X = a + (-sin phi - i cos phi) b = a + (w.y - i w.x) b
        const cx<T> f = {w.y, -w.x};
        spec[(size_t)(L0 + l) * K3h + k] = a + cmul(f, b);
    }
}

// inverse: K3/2+1 complex -> K3 reals, unnormalised (tools/fft_model.py c2r)
template <typename T>
__global__ void __launch_bounds__(256)
fft_z_inv_kernel(FftPlan P, int TL, int LS, int nlines, int K3, const cx<T>* __restrict__ spec, T* __restrict__ mesh,
                 const cx<T>* __restrict__ gtw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* A = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* Bf = A + TL * LS;
    cx<T>* stw = Bf + TL * LS;
    const int M = K3 / 2, K3h = M + 1;
    const int L0 = blockIdx.x * TL;
    const int nl = min(TL, nlines - L0);
    load_twiddles(stw, gtw, K3);
    // stage the half spectrum in B (needs M+1 <= LS: LS = M|1 >= M+... ensured by the host: LS >= M+1)
    for (int e = threadIdx.x; e < nl * K3h; e += blockDim.x) {
        const int l = e / K3h, k = e - l * K3h;
        Bf[l * LS + k] = spec[(size_t)(L0 + l) * K3h + k];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nl * M; e += blockDim.x) {
        const int l = e / M, k = e - l * M;
        const cx<T> xk = Bf[l * LS + k];
        cx<T> xc = Bf[l * LS + M - k];
        xc.y = -xc.y;
        const cx<T> s = xk + xc, d = xk - xc;
        const cx<T> w = stw[k];                                   // conj gives (cos phi, +sin phi)
        // Z = s + i (cos phi + i sin phi) d = s + (-sin phi + i cos phi) d = s + (w.y + i w.x) d
        const cx<T> f = {w.y, w.x};
        A[l * LS + k] = s + cmul(f, d);
    }
    __syncthreads();
    cx<T>* z = fft_lines<T, -1>(A, Bf, P, nl, LS, stw, 2);
    const T* zr = reinterpret_cast<const T*>(z);
    for (int e = threadIdx.x; e < nl * K3; e += blockDim.x) {
        const int l = e / K3, n = e - l * K3;
        mesh[(size_t)(L0 + l) * K3 + n] = zr[2 * l * LS + n];
    }
}

// ------------------------------------------------------------------------------------------ strided passes (Y, X)
struct StrideGeom {
    int n_outer;          // Y: K1 ; X: 1
    int n_inner;          // Y: K3h ; X: K2*K3h   (contiguous index the TL lines of a block walk)
    size_t outer_stride;  // Y: K2*K3h ; X: 0
    size_t line_stride;   // Y: K3h ; X: K2*K3h   (distance between consecutive points of one line)
    int tiles;            // ceil(n_inner / TL)
};

template <typename T>
__device__ __forceinline__ void load_tile(cx<T>* A, const cx<T>* __restrict__ g, int N, int nl, int TL, int LS, size_t stride) {
    for (int e = threadIdx.x; e < N * TL; e += blockDim.x) {
        const int pos = e / TL, l = e - pos * TL;
        if (l < nl) A[l * LS + pos] = g[(size_t)pos * stride + l];
    }
}
template <typename T>
__device__ __forceinline__ void store_tile(const cx<T>* A, cx<T>* __restrict__ g, int N, int nl, int TL, int LS, size_t stride) {
    for (int e = threadIdx.x; e < N * TL; e += blockDim.x) {
        const int pos = e / TL, l = e - pos * TL;
        if (l < nl) g[(size_t)pos * stride + l] = A[l * LS + pos];
    }
}

template <typename T, int SIGN>
__global__ void __launch_bounds__(256)
fft_strided_kernel(FftPlan P, int TL, int LS, StrideGeom g, cx<T>* __restrict__ spec, const cx<T>* __restrict__ gtw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* A = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* Bf = A + TL * LS;
    cx<T>* stw = Bf + TL * LS;
    const int o = blockIdx.x / g.tiles, t = blockIdx.x - o * g.tiles;
    const int c0 = t * TL;
    const int nl = min(TL, g.n_inner - c0);
    cx<T>* base = spec + (size_t)o * g.outer_stride + c0;
    load_twiddles(stw, gtw, P.N);
    load_tile(A, base, P.N, nl, TL, LS, g.line_stride);
    __syncthreads();
    cx<T>* R = fft_lines<T, SIGN>(A, Bf, P, nl, LS, stw, 1);
    store_tile(R, base, P.N, nl, TL, LS, g.line_stride);
}

// X-forward, multiply by 2*scale*C_k/theta_k^2 with energy (+virial) accumulation, X-inverse: the
// x-transformed spectrum only ever lives in shared memory (admp/recip.py:410-426 fused with its adjoint)
template <typename T>
__global__ void __launch_bounds__(256)
fft_x_conv_kernel(FftPlan P, int TL, int LS, StrideGeom g, const BoxInfo* __restrict__ Bp, T kappa, int kind, ConvTables tb,
                  cx<T>* __restrict__ spec, const cx<T>* __restrict__ gtw, double* __restrict__ scalars, int want_vir) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[7 * 8];
    cx<T>* A = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* Bf = A + TL * LS;
    cx<T>* stw = Bf + TL * LS;
    const BoxInfo& B = *Bp;
    const int c0 = blockIdx.x * TL;
    const int nl = min(TL, g.n_inner - c0);
    cx<T>* base = spec + c0;
    load_twiddles(stw, gtw, P.N);
    load_tile(A, base, P.N, nl, TL, LS, g.line_stride);
    __syncthreads();
    cx<T>* R = fft_lines<T, 1>(A, Bf, P, nl, LS, stw, 1);
    cx<T>* O = (R == A) ? Bf : A;
    // influence function on the tile: element (pos = i1, line l -> flattened (i2, i3))
    const int K3 = B.K[2], K3h = K3 / 2 + 1;
    const double scale = (kind == ADMP_CK_COULOMB) ? ADMP_DIEL : 1.0;
    const bool ortho = *tb.ortho != 0;
    const double kap = (double)kappa;
    double acc_e = 0.0, acc_t[6] = {0, 0, 0, 0, 0, 0};
    for (int e = threadIdx.x; e < P.N * nl; e += blockDim.x) {
        const int l = e / P.N, i1 = e - l * P.N;
        const int c = c0 + l;
        const int i2 = c / K3h, i3 = c - i2 * K3h;
        cx<T> s = R[l * LS + i1];
        const double s2 = (double)s.x * s.x + (double)s.y * s.y;
        const bool single = (i3 == 0) || (2 * i3 == K3);
        double gk;
        if (want_vir) {
            const Influence f = influence<true>(B, tb, ortho, kap, kind, i1, i2, i3);
            virial_terms(B, f.kv, i1, i2, i3, single, f.dg * s2, acc_t);
            gk = f.g;
        } else {
            gk = influence<false>(B, tb, ortho, kap, kind, i1, i2, i3).g;
        }
        acc_e += (single ? 1.0 : 2.0) * gk * s2;
        const T gg = (T)(2.0 * scale * gk);
        s.x *= gg; s.y *= gg;
        R[l * LS + i1] = s;
    }
    __syncthreads();
    cx<T>* R2 = fft_lines<T, -1>(R, O, P, nl, LS, stw, 1);
    store_tile(R2, base, P.N, nl, TL, LS, g.line_stride);
    double e1[1] = {acc_e * scale};
    block_accumulate<1>(e1, red, scalars + ADMP_S_E_RECIP);
    if (want_vir) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc_t[k] *= scale;
        block_accumulate<6>(acc_t, red, scalars + ADMP_S_TK);
    }
}

#include "fft_fast.cuh"

// ------------------------------------------------------------------------------------------ host side
bool fft_factorize(int n, FftPlan& P) {
    P.N = n;
    P.nst = 0;
    const int odd[5] = {13, 11, 7, 5, 3};
    for (int r : odd)
        while (n % r == 0) { if (P.nst >= 12) return false; P.radix[P.nst++] = r; n /= r; }
    while (n % 4 == 0) { if (P.nst >= 12) return false; P.radix[P.nst++] = 4; n /= 4; }
    while (n % 2 == 0) { if (P.nst >= 12) return false; P.radix[P.nst++] = 2; n /= 2; }
    return n == 1 && P.nst > 0;
}

struct FftDimCfg {
    FftPlan P; int TL, LS; size_t smem;          // generic Stockham kernels
    bool fast;                                   // pipelined register-blocked kernels (fft_fast.cuh)
    FastOps ops;
};

// zfwd: configuration of the FORWARD Z pass alone - it wants 8-line tiles (line-fastest thread mapping, fft_fast.cuh), the inverse Z
// pass the narrow ones; only the forward kernel of the entry has to fit
static bool dim_cfg(int N, int tw_len, size_t esz, size_t smem_cap, int min_ls, bool zpass, bool allow_fast, FftDimCfg& c, bool xpass = false,
                    bool zfwd = false) {
    if (!fft_factorize(N, c.P)) return false;
    c.LS = N | 1;
    if (c.LS < min_ls) c.LS = min_ls | 1;
    bool ok = false;
    for (int TL = 8; TL >= 1; TL >>= 1) {
        const size_t need = (size_t)2 * TL * c.LS * 2 * esz + (size_t)tw_len * 2 * esz;
        if (need <= smem_cap) { c.TL = TL; c.smem = need; ok = true; break; }
    }
    if (!ok) return false;
    c.fast = false;
    if (allow_fast) {
        // two tile widths exist for the larger sizes: strided passes default to the wide tile (longer contiguous
        // global segments), the contiguous Z passes to the narrow one (more resident blocks); measured on B200
        const char* e = getenv("ADMP_FFT_WIDE");
        bool wide = e ? atoi(e) > 0 : (!zpass || zfwd);
        if (xpass) { const char* ex = getenv("ADMP_FFT_XWIDE"); if (ex) wide = atoi(ex) > 0; }
        int force = -1;
        if (xpass && N == 616) force = 9;          // X pass of 616 points: one 448-thread block on 8-line tiles (6.73 -> 6.13 ms at 616x1232x1232)
        if (!zpass) { const char* ef = getenv(xpass ? "ADMP_FFT_XCFG" : "ADMP_FFT_YCFG"); if (ef) force = atoi(ef); }
        if (zfwd) {
            if (N == 616) force = 11;              // 8-line tiles in one 448-thread block
            const char* ef = getenv("ADMP_FFT_ZFCFG");
            if (ef) force = atoi(ef);
        }
        c.fast = esz == 8 ? fast_lookup<double>(N, wide, c.ops, force) : fast_lookup<float>(N, wide, c.ops, force);
        if (c.fast) {
            c.ops.prepare(c.ops);
            const bool usable = zfwd ? (c.ops.occ[3] > 0) : zpass ? (c.ops.occ[3] > 0 && c.ops.occ[4] > 0) : (c.ops.occ[0] > 0 && c.ops.occ[1] > 0 && c.ops.occ[2] > 0 && c.ops.occ[5] > 0);
            if (!usable) c.fast = false;
        }
    }
    return true;
}

struct Fft3dImpl {
    int K[3];
    size_t esz;
    int n_sm;
    FftDimCfg z, y, x;
    FftDimCfg zf;    // forward Z pass (8-line tiles where they exist); z serves the inverse pass and the generic fallback
    void* tw[3];     // device twiddle tables: exp(-2 pi i m / K_d), m < K_d
};

template <typename T>
static cudaError_t set_smem_attr(const Fft3dImpl* f) {
    cudaError_t e;
    const size_t zs = f->z.smem, ys = f->y.smem, xs = f->x.smem;
    if ((e = cudaFuncSetAttribute(fft_z_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_z_inv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs)) != cudaSuccess) return e;
    const size_t m = ys > xs ? ys : xs;
    if ((e = cudaFuncSetAttribute(fft_strided_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_strided_kernel<T, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(fft_x_conv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xs)) != cudaSuccess) return e;
    return cudaSuccess;
}

Fft3d* fft3d_create(int K1, int K2, int K3, int dtype, const char** why) {
    static const char* msg_odd = "K3 is odd";
    static const char* msg_fac = "a mesh dimension has a prime factor outside {2,3,5,7,11,13} or does not fit in shared memory";
    static const char* msg_cuda = "CUDA error while preparing the FFT";
    if (K3 % 2) { *why = msg_odd; return nullptr; }
    Fft3dImpl* f = new Fft3dImpl();
    f->K[0] = K1; f->K[1] = K2; f->K[2] = K3;
    f->esz = dtype == ADMP_F64 ? 8 : 4;
    int dev = 0;
    cudaGetDevice(&dev);
    f->n_sm = 148;
    cudaDeviceGetAttribute(&f->n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (upload_trig_tables() != cudaSuccess) { cudaGetLastError(); *why = msg_cuda; delete f; return nullptr; }
    const size_t cap = 200 * 1024;
    const int M = K3 / 2;
    const char* env = getenv("ADMP_FFT");
    const bool allow_fast = !(env && strcmp(env, "generic") == 0);
    if (!dim_cfg(M, K3, f->esz, cap, M + 1, true, allow_fast, f->z) || !dim_cfg(K2, K2, f->esz, cap, 0, false, allow_fast, f->y) ||
        !dim_cfg(K1, K1, f->esz, cap, 0, false, allow_fast, f->x, true)) {
        *why = msg_fac;
        delete f;
        return nullptr;
    }
    f->zf = f->z;
    if (f->z.fast) {
        FftDimCfg zf;
        if (dim_cfg(M, K3, f->esz, cap, M + 1, true, allow_fast, zf, false, true) && zf.fast) f->zf = zf;
    }
    cudaError_t e = dtype == ADMP_F64 ? set_smem_attr<double>(f) : set_smem_attr<float>(f);
    if (e != cudaSuccess) { *why = msg_cuda; delete f; return nullptr; }
    for (int d = 0; d < 3; ++d) {
        const int n = f->K[d];
        std::vector<double> h(2 * (size_t)n);
        for (int m = 0; m < n; ++m) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
            h[2 * m] = (double)cosl(a);
            h[2 * m + 1] = (double)sinl(a);
        }
        if (cudaMalloc(&f->tw[d], 2 * (size_t)n * f->esz) != cudaSuccess) { *why = msg_cuda; delete f; return nullptr; }
        if (dtype == ADMP_F64) cudaMemcpy(f->tw[d], h.data(), 2 * (size_t)n * 8, cudaMemcpyHostToDevice);
        else {
            std::vector<float> hf(h.begin(), h.end());
            cudaMemcpy(f->tw[d], hf.data(), 2 * (size_t)n * 4, cudaMemcpyHostToDevice);
        }
    }
    return reinterpret_cast<Fft3d*>(f);
}

void fft3d_destroy(Fft3d* p) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    if (!f) return;
    for (int d = 0; d < 3; ++d) if (f->tw[d]) cudaFree(f->tw[d]);
    delete f;
}

// ---- TMA tensor maps of the strided tiles (float64): dims {2*n_inner doubles, N positions, n_outer planes}, box
// {2*TL, box_rows, 1}, no swizzle, out-of-range columns zero-filled. The encoder comes from the driver at run time
// (cudaGetDriverEntryPoint: no link against libcuda); ADMP_FFT_TMA=0, a missing encoder or a failed encode select cp.async.
static PFN_cuTensorMapEncodeTiled tmap_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = [] {
        const char* e = getenv("ADMP_FFT_TMA");
        if (e && atoi(e) == 0) return (PFN_cuTensorMapEncodeTiled) nullptr;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (PFN_cuTensorMapEncodeTiled) nullptr;
        }
        return (PFN_cuTensorMapEncodeTiled)p;
    }();
    return fn;
}
static bool make_tmap(CUtensorMap* tm, void* base, const StrideGeom& g, int N, int TL, int box_rows) {
    PFN_cuTensorMapEncodeTiled enc = tmap_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {2 * (cuuint64_t)g.n_inner, (cuuint64_t)N, (cuuint64_t)g.n_outer};
    const cuuint64_t strides[2] = {(cuuint64_t)g.line_stride * 16, (cuuint64_t)(g.n_outer > 1 ? g.outer_stride : (size_t)N * g.line_stride) * 16};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * TL), (cuuint32_t)box_rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static StrideGeom geom_y(const Fft3dImpl* f, int TL) {
    const int K1 = f->K[0], K2 = f->K[1], K3h = f->K[2] / 2 + 1;
    return {K1, K3h, (size_t)K2 * K3h, (size_t)K3h, (K3h + TL - 1) / TL};
}
static StrideGeom geom_x(const Fft3dImpl* f, int TL) {
    const int inner = f->K[1] * (f->K[2] / 2 + 1);
    return {1, inner, 0, (size_t)inner, (inner + TL - 1) / TL};
}
// persistent grid: at most `occ` resident blocks per SM, tiles dealt round-robin; the grid is shrunk so that
// every block gets the same number of tiles (no block does one tile more than the rest: on the L2-resident
// 154^3 mesh a pass is only 2-3 tiles per block, so an uneven deal costs a third of the pass)
static int persistent_grid(const Fft3dImpl* f, int occ, int ntiles) {
    // default on: with four evaluations in flight the slots a shrunk grid leaves free take the other evaluations' blocks
    // (C2 469 -> 477 evals/s, twice each on one lease; a pass timed alone is 1-4 % slower); ADMP_FFT_BALANCED=0: all resident slots
    static const bool balanced = [] { const char* e = getenv("ADMP_FFT_BALANCED"); return !(e && atoi(e) == 0); }();
    // ADMP_FFT_GRID_DIV = d: a pass only asks for 1/d of the resident-block slots, so that the passes of d independent
    // evaluations in flight on different streams run side by side instead of queueing behind each other's persistent blocks
    static const int div = [] { const char* e = getenv("ADMP_FFT_GRID_DIV"); const int v = e ? atoi(e) : 1; return v > 0 ? v : 1; }();
    long long cap = (long long)f->n_sm * (occ > 0 ? occ : 1) / div;
    if (cap < 1) cap = 1;
    if (ntiles <= cap) return ntiles;
    if (!balanced) return (int)cap;
    const long long waves = (ntiles + cap - 1) / cap;
    return (int)((ntiles + waves - 1) / waves);
}

// x0 / nx: restrict the pass to the planes [x0, x0 + nx) of the buffers (x-slab decomposition); nx < 0 = all
template <typename T>
static void run_z(Fft3dImpl* f, cudaStream_t st, void* mesh, void* spec, int sign, int x0 = 0, int nx = -1, int zld = 0) {
    const int K3 = f->K[2];
    if (nx < 0) nx = f->K[0];
    const int nlines = nx * f->K[1];
    if (zld <= 0) zld = K3;                   // reals per mesh line: K3, or 2 (K3/2 + 1) when the mesh lives in the spectrum buffer
    mesh = (char*)mesh + (size_t)x0 * f->K[1] * zld * sizeof(T);
    spec = (char*)spec + (size_t)x0 * f->K[1] * (K3 / 2 + 1) * sizeof(cx<T>);
    const FftDimCfg& c = (sign > 0) ? f->zf : f->z;
    const cx<T>* tw = (const cx<T>*)f->tw[2];
    if (c.fast) {
        const int ntiles = (nlines + c.ops.zTL - 1) / c.ops.zTL;
        if (sign > 0) c.ops.zfwd(st, nlines, ntiles, persistent_grid(f, c.ops.occ[3], ntiles), mesh, spec, tw, zld / 2);
        else c.ops.zinv(st, nlines, ntiles, persistent_grid(f, c.ops.occ[4], ntiles), spec, mesh, tw, zld / 2);
        return;
    }
    const int grid = (nlines + c.TL - 1) / c.TL;
    if (sign > 0) fft_z_fwd_kernel<T><<<grid, 256, c.smem, st>>>(c.P, c.TL, c.LS, nlines, K3, (const T*)mesh, (cx<T>*)spec, tw);
    else fft_z_inv_kernel<T><<<grid, 256, c.smem, st>>>(c.P, c.TL, c.LS, nlines, K3, (const cx<T>*)spec, (T*)mesh, tw);
}

template <typename T>
static void run_strided(Fft3dImpl* f, cudaStream_t st, void* spec, int dim, int sign, int x0 = 0, int nx = -1) {
    const FftDimCfg& c = dim == 1 ? f->y : f->x;
    const cx<T>* tw = (const cx<T>*)f->tw[dim == 1 ? 1 : 0];
    if (c.fast) {
        StrideGeom g = dim == 1 ? geom_y(f, c.ops.TL) : geom_x(f, c.ops.TL);
        if (dim == 1 && nx >= 0) {
            g.n_outer = nx;
            spec = (char*)spec + (size_t)x0 * g.outer_stride * sizeof(cx<T>);
        }
        const int ntiles = g.n_outer * g.tiles;
        CUtensorMap tm;
        if (sizeof(T) == 8 && c.ops.occ_tma[sign > 0 ? 0 : 1] > 0 && make_tmap(&tm, spec, g, c.ops.N, c.ops.TL, c.ops.box_rows)) {
            c.ops.strided_tma(st, sign, g, ntiles, persistent_grid(f, c.ops.occ_tma[sign > 0 ? 0 : 1], ntiles), spec, tw, tm);
            return;
        }
        c.ops.strided(st, sign, g, ntiles, persistent_grid(f, c.ops.occ[sign > 0 ? 0 : 1], ntiles), spec, tw);
        return;
    }
    const StrideGeom g = dim == 1 ? geom_y(f, c.TL) : geom_x(f, c.TL);
    const int grid = g.n_outer * g.tiles;
    if (sign > 0) fft_strided_kernel<T, 1><<<grid, 256, c.smem, st>>>(c.P, c.TL, c.LS, g, (cx<T>*)spec, tw);
    else fft_strided_kernel<T, -1><<<grid, 256, c.smem, st>>>(c.P, c.TL, c.LS, g, (cx<T>*)spec, tw);
}

template <typename T>
static void run_x_conv(Fft3dImpl* f, cudaStream_t st, void* spec, const BoxInfo* B, double kappa, int kind, const ConvTables& tb,
                       double* scalars, int want_vir) {
    const FftDimCfg& c = f->x;
    const cx<T>* tw = (const cx<T>*)f->tw[0];
    if (c.fast) {
        const StrideGeom g = geom_x(f, c.ops.TL);
        const bool quick = (kind == ADMP_CK_COULOMB && !want_vir);
        const bool qv = (kind == ADMP_CK_COULOMB && want_vir);        // quick + virial variant when it fits (occ_qv > 0)
        const int occ_t = (qv && c.ops.occ_qv[1] > 0) ? c.ops.occ_qv[1] : c.ops.occ_tma[quick ? 2 : 3];
        const int occ_c = (qv && c.ops.occ_qv[0] > 0) ? c.ops.occ_qv[0] : c.ops.occ[quick ? 2 : 5];
        CUtensorMap tm;
        if (sizeof(T) == 8 && occ_t > 0 && c.ops.occ_tma[3] > 0 && make_tmap(&tm, spec, g, c.ops.N, c.ops.TL, c.ops.box_rows)) {
            c.ops.xconv_tma(st, g, 0, g.tiles, persistent_grid(f, occ_t, g.tiles), B, kappa, kind, tb, spec, tw, scalars, want_vir, tm);
            return;
        }
        c.ops.xconv(st, g, 0, g.tiles, persistent_grid(f, occ_c, g.tiles), B, kappa, kind, tb, spec, tw, scalars, want_vir);
        return;
    }
    const StrideGeom g = geom_x(f, c.TL);
    fft_x_conv_kernel<T><<<g.tiles, 256, c.smem, st>>>(c.P, c.TL, c.LS, g, B, (T)kappa, kind, tb, (cx<T>*)spec, tw, scalars, want_vir);
}

// ---- x-slab decomposed round trip over several GPUs (each rank owns K1/n consecutive x planes of mesh and
// spectrum; buffers keep the full-size indexing). phase 0: Z-forward + Y-forward on the own planes;
// phase 1: fused X pass (forward, influence function, inverse) on this rank's share of the (y, kz) columns,
// reading and writing all ranks' planes through the peer table; phase 2: Y-inverse + Z-inverse on the own planes.
// The caller places a cross-rank barrier between the phases.
bool fft3d_slab_supported(const Fft3d* p) {
    const Fft3dImpl* f = reinterpret_cast<const Fft3dImpl*>(p);
    return f && f->z.fast && f->y.fast && f->x.fast && f->x.ops.occ_peer[0] > 0 && f->x.ops.occ_peer[1] > 0;
}
// phase 1 pipeline (SlabAux non-null, n > 1): the peers' planes of this rank's columns are pulled into the own
// spectrum buffer chunk by chunk with strided peer copies on a copy stream (DMA engines over NVLink, large
// contiguous rows), the fused X pass of chunk k runs as soon as its pull has landed and stores its results
// straight to the owners' planes (peer stores, fire and forget): pull(k+1) | transform(k) | push(k) overlap.
// ADMP_SLAB_PULL=0 selects the fully in-kernel variant (peer loads through cp.async as well).
template <typename T>
static void slab_phase(Fft3dImpl* f, cudaStream_t st, int phase, int rank, void* mesh, void* spec, const PeerTab& peers, const BoxInfo* B,
                       double kappa, int kind, const ConvTables& tb, double* scalars, int want_vir, const SlabAux* aux) {
    const int x0 = peers.start[rank], nx = peers.planes(rank);
    if (phase == 0) {
        run_z<T>(f, st, mesh, spec, 1, x0, nx);
        run_strided<T>(f, st, spec, 1, 1, x0, nx);
    } else if (phase == 1) {
        const FftDimCfg& c = f->x;
        const StrideGeom g = geom_x(f, c.ops.TL);
        const int t0 = (int)((long long)g.tiles * rank / peers.n), t1 = (int)((long long)g.tiles * (rank + 1) / peers.n);
        if (t1 <= t0) return;
        const bool quick = kind == ADMP_CK_COULOMB && !want_vir;
        const int occ = c.ops.occ_peer[quick ? 0 : 1];
        static const bool pull = [] { const char* e = getenv("ADMP_SLAB_PULL"); return !(e && atoi(e) == 0); }();
        if (!pull || !aux || peers.n == 1) {
            c.ops.xconv_peer(st, g, t0, t1, persistent_grid(f, occ, t1 - t0), B, kappa, kind, tb, spec, f->tw[0], scalars, want_vir, peers, 0);
            return;
        }
        static const int want_chunks = [] { const char* e = getenv("ADMP_SLAB_CHUNKS"); return e ? atoi(e) : 4; }();
        static const int debug_skip = [] { const char* e = getenv("ADMP_SLAB_SKIP"); return e ? atoi(e) : 0; }();   // 1: no copies, 2: no kernels (timing only)
        const int nchunk = std::max(1, std::min(SLAB_CHUNKS, std::min(want_chunks, t1 - t0)));
        const size_t pitch = (size_t)g.n_inner * sizeof(cx<T>);
        // copies of one chunk: one per peer (two row halves per peer when there is a single peer), dealt over the
        // copy streams; staggered source order: at step s every rank pulls from rank + s, so each source feeds
        // exactly one reader at a time
        // measured on B200 / NVSwitch: ONE copy stream is fastest (P = 2: 7.3 vs 7.8 ms per pass, P = 4: 5.5 vs 6.2 ms
        // with four); ADMP_SLAB_STREAMS spreads the pulls of a chunk over more streams for experiments
        static const int want_streams = [] { const char* e = getenv("ADMP_SLAB_STREAMS"); return e ? atoi(e) : 1; }();
        const int halves = (peers.n == 2 && want_streams > 1) ? 2 : 1;
        const int nstreams = std::max(1, std::min(std::min(SLAB_STREAMS, want_streams), (peers.n - 1) * halves));
        cudaEventRecord(aux->fork, st);
        for (int s = 0; s < nstreams; ++s) cudaStreamWaitEvent(aux->copy_stream[s], aux->fork, 0);
        for (int k = 0; k < nchunk; ++k) {
            const int a = t0 + (int)((long long)(t1 - t0) * k / nchunk), b = t0 + (int)((long long)(t1 - t0) * (k + 1) / nchunk);
            const size_t col0 = (size_t)a * c.ops.TL, col1 = std::min((size_t)b * c.ops.TL, (size_t)g.n_inner);
            int item = 0;
            for (int s = 1; s < peers.n; ++s) {
                const int q = (rank + s) % peers.n;
                for (int h = 0; h < halves; ++h, ++item) {
                    if (col1 <= col0 || debug_skip == 1) continue;
                    const int r0 = peers.planes(q) * h / halves, r1 = peers.planes(q) * (h + 1) / halves;
                    const size_t off = (((size_t)peers.start[q] + r0) * g.n_inner + col0) * sizeof(cx<T>);
                    cudaMemcpy2DAsync((char*)spec + off, pitch, (const char*)peers.base[q] + off, pitch, (col1 - col0) * sizeof(cx<T>),
                                      (size_t)(r1 - r0), cudaMemcpyDefault, aux->copy_stream[item % nstreams]);
                }
            }
            for (int s = 0; s < nstreams; ++s) cudaEventRecord(aux->chunk[k][s], aux->copy_stream[s]);
        }
        // results go back to the owners as peer stores from inside the kernel (default) or, with ADMP_SLAB_PUSH=dma, as
        // strided peer copies on a second copy stream behind a purely local transform (pull(k+1) | transform(k) |
        // push(k-1) on three engines). Measured on B200 / NVSwitch: both are bound by the NVLink ingress of a rank
        // (its pulls + its peers' pushes, ~500 GB/s aggregate): 5.7 (kernel) vs 6.2 ms (dma) per pass at 4 ranks.
        static const bool dma_push = [] { const char* e = getenv("ADMP_SLAB_PUSH"); return e && strcmp(e, "dma") == 0; }();
        cudaStream_t push_stream = aux->copy_stream[SLAB_STREAMS - 1];
        const int occ_local = c.ops.occ[quick ? 2 : 5];
        for (int k = 0; k < nchunk; ++k) {
            const int a = t0 + (int)((long long)(t1 - t0) * k / nchunk), b = t0 + (int)((long long)(t1 - t0) * (k + 1) / nchunk);
            for (int s = 0; s < nstreams; ++s) cudaStreamWaitEvent(st, aux->chunk[k][s], 0);
            if (b <= a || debug_skip == 2) continue;
            if (!dma_push) {
                c.ops.xconv_peer(st, g, a, b, persistent_grid(f, occ, b - a), B, kappa, kind, tb, spec, f->tw[0], scalars, want_vir, peers, 1);
                continue;
            }
            c.ops.xconv(st, g, a, b, persistent_grid(f, occ_local, b - a), B, kappa, kind, tb, spec, f->tw[0], scalars, want_vir);
            cudaEventRecord(aux->done[k], st);
            cudaStreamWaitEvent(push_stream, aux->done[k], 0);
            const size_t col0 = (size_t)a * c.ops.TL, col1 = std::min((size_t)b * c.ops.TL, (size_t)g.n_inner);
            for (int s = 1; s < peers.n; ++s) {
                const int q = (rank + s) % peers.n;
                const size_t off = ((size_t)peers.start[q] * g.n_inner + col0) * sizeof(cx<T>);
                cudaMemcpy2DAsync((char*)peers.base[q] + off, pitch, (const char*)spec + off, pitch, (col1 - col0) * sizeof(cx<T>),
                                  (size_t)peers.planes(q), cudaMemcpyDefault, push_stream);
            }
        }
        if (dma_push) {
            cudaEventRecord(aux->pushed, push_stream);
            cudaStreamWaitEvent(st, aux->pushed, 0);
        }
    } else {
        run_strided<T>(f, st, spec, 1, -1, x0, nx);
        run_z<T>(f, st, mesh, spec, -1, x0, nx);
    }
}
void fft3d_slab_phase(Fft3d* p, cudaStream_t st, int phase, int rank, void* mesh, void* spec, const PeerTab& peers, const BoxInfo* B,
                      double kappa, int kind, const ConvTables& tb, double* scalars, int want_vir, const SlabAux* aux) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    if (f->esz == 8) slab_phase<double>(f, st, phase, rank, mesh, spec, peers, B, kappa, kind, tb, scalars, want_vir, aux);
    else slab_phase<float>(f, st, phase, rank, mesh, spec, peers, B, kappa, kind, tb, scalars, want_vir, aux);
}

// plain transforms (same conventions as cuFFT D2Z / Z2D: unnormalised)
void fft3d_forward(Fft3d* p, cudaStream_t st, const void* mesh, void* spec) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    if (f->esz == 8) { run_z<double>(f, st, const_cast<void*>(mesh), spec, 1); run_strided<double>(f, st, spec, 1, 1); run_strided<double>(f, st, spec, 0, 1); }
    else { run_z<float>(f, st, const_cast<void*>(mesh), spec, 1); run_strided<float>(f, st, spec, 1, 1); run_strided<float>(f, st, spec, 0, 1); }
}
void fft3d_inverse(Fft3d* p, cudaStream_t st, void* spec, void* mesh) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    if (f->esz == 8) { run_strided<double>(f, st, spec, 0, -1); run_strided<double>(f, st, spec, 1, -1); run_z<double>(f, st, mesh, spec, -1); }
    else { run_strided<float>(f, st, spec, 0, -1); run_strided<float>(f, st, spec, 1, -1); run_z<float>(f, st, mesh, spec, -1); }
}

// a single pass of the five (0: Z-forward, 1: Y-forward, 2: fused X, 3: Y-inverse, 4: Z-inverse): profiling / roofline timing
void fft3d_single_pass(Fft3d* p, cudaStream_t st, int which, void* mesh, void* spec, const BoxInfo* B, double kappa, int kind,
                       const ConvTables& tb, double* scalars) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    const bool d = f->esz == 8;
    switch (which) {
        case 0: d ? run_z<double>(f, st, mesh, spec, 1) : run_z<float>(f, st, mesh, spec, 1); break;
        case 1: d ? run_strided<double>(f, st, spec, 1, 1) : run_strided<float>(f, st, spec, 1, 1); break;
        case 2: d ? run_x_conv<double>(f, st, spec, B, kappa, kind, tb, scalars, 0) : run_x_conv<float>(f, st, spec, B, kappa, kind, tb, scalars, 0); break;
        case 3: d ? run_strided<double>(f, st, spec, 1, -1) : run_strided<float>(f, st, spec, 1, -1); break;
        default: d ? run_z<double>(f, st, mesh, spec, -1) : run_z<float>(f, st, mesh, spec, -1); break;
    }
}

// mesh -> phi = dE/dmesh in place of the mesh, energy (+virial sums) accumulated: 5 passes
// mesh_out: where the inverse Z pass writes the potential (nullptr: over the input mesh); after_zfwd (optional): recorded
// once the forward Z pass has consumed `mesh`, so the caller can recycle it while the spectrum passes run
void fft3d_convolve_roundtrip(Fft3d* p, cudaStream_t st, void* mesh, void* spec, const BoxInfo* B, double kappa, int kind,
                              const ConvTables& tb, double* scalars, int want_vir, void* mesh_out, cudaEvent_t after_zfwd, int zld) {
    Fft3dImpl* f = reinterpret_cast<Fft3dImpl*>(p);
    if (mesh_out == nullptr) mesh_out = mesh;
    if (f->esz == 8) {
        run_z<double>(f, st, mesh, spec, 1, 0, -1, zld);
        run_strided<double>(f, st, spec, 1, 1);
        if (after_zfwd) cudaEventRecord(after_zfwd, st);      // recorded in front of the X pass: FP64-bound, leaves DRAM idle
        run_x_conv<double>(f, st, spec, B, kappa, kind, tb, scalars, want_vir);
        run_strided<double>(f, st, spec, 1, -1);
        run_z<double>(f, st, mesh_out, spec, -1, 0, -1, zld);
    } else {
        run_z<float>(f, st, mesh, spec, 1, 0, -1, zld);
        run_strided<float>(f, st, spec, 1, 1);
        if (after_zfwd) cudaEventRecord(after_zfwd, st);
        run_x_conv<float>(f, st, spec, B, kappa, kind, tb, scalars, want_vir);
        run_strided<float>(f, st, spec, 1, -1);
        run_z<float>(f, st, mesh_out, spec, -1, 0, -1, zld);
    }
}

// the real mesh may live in the spectrum buffer (line l of K3 reals at the start of spectrum line l, zld = 2 (K3/2 + 1)): only the
// tile-pipelined Z passes read a whole tile into shared memory before they write it
bool fft3d_inplace_supported(const Fft3d* p) {
    const Fft3dImpl* f = reinterpret_cast<const Fft3dImpl*>(p);
    return f->zf.fast && f->z.fast;
}

}  // namespace admp
