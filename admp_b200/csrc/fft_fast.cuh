// Fragment of fft.cu (included inside namespace admp, after Dft<> / cx<> / StrideGeom):
// the pipelined line-FFT kernels for the reference's mesh family K = 154 * 2^n (and the halves the
// packed real transform needs): N = R1*R2*R3 in {77, 154, 308, 616, 1232}.
//
// Design (B200, FP64: the passes need ~30-70 FP64 instructions per 16-byte point, so both the
// HBM stream and the FP64 pipe have to stay busy at the same time):
//   * persistent blocks, one tile (TL lines) per loop iteration, tiles round-robin over the grid;
//   * two tile buffers per block: I (cp.async / LDGSTS target) and A (work). Stage 1 moves the tile
//     I -> A; as soon as it is complete the NEXT tile is fetched into I with cp.async while the
//     remaining stages run (a middle stage of a three-stage transform exchanges in place in A:
//     shared -> registers, DFT, barrier, registers -> shared), so no thread waits on a global
//     load inside the butterfly code and five 112-thread blocks fit on an SM;
//   * strided passes (Y, X) keep the tile position-major in shared memory ([pos][TL], TL
//     consecutive lines = one contiguous global segment), thread = (line fastest, butterfly):
//     every shared-memory access of a quarter warp is one contiguous 128-byte row - no bank
//     conflicts, no padding, and the cp.async chunks map 1:1 onto global segments;
//   * butterflies are register blocked (one radix-R DFT per thread per step, JT threads per
//     line looping over the N/R butterflies of a stage), twiddles come from compact per-stage
//     shared-memory tables laid out [t][k] (conflict-free), all offsets are compile-time constants;
//   * the X pass is fused with the influence function: X-forward, C_k/theta_k^2 scaling with the
//     energy (+ virial) sums, X-inverse - the x-transformed spectrum never leaves the SM.

// ------------------------------------------------------------------------------------------ cp.async
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TMA bulk copies (sm_90+)
// Tile loads of the float64 passes go through the TMA unit: `cp.async.bulk.shared.global` (SASS: UBLKCP) moves a whole
// contiguous run - a 128-byte row of a strided tile, a full line of a Z tile - per instruction and signals an mbarrier with the
// bytes it delivered; no per-thread 16-byte LDGSTS, no per-thread address arithmetic in registers, no cp.async group to drain.
// One mbarrier per block (a single tile is in flight), phase parity flips per tile. The float passes keep cp.async (their
// partial rows are not multiples of 16 bytes).
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// the generic-proxy reads of the tile buffer are ordered before the async-proxy writes that refill it
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();          // a lost transaction must fail loudly, not hang the device
    }
}
// one 3-D tensor-map copy (SASS: UTMALDG): box {2*TL doubles, BOXN positions, 1 plane} of the spectrum -> shared memory
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(tm), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// rows of a strided tile per tensor-map box: the largest divisor of N within the 256-element box limit whose boxes start
// on 128-byte shared-memory boundaries (rows are TL complex doubles = TL*16 bytes); 0 = no such divisor (cp.async is used)
constexpr int tma_box_rows(int N, int TL) {
    int b = 0;
    for (int d = 1; d <= 256; ++d)
        if (N % d == 0 && (d == N || (d * TL) % 8 == 0)) b = d;
    return b;
}
template <typename T> struct UseBulk { static constexpr bool value = sizeof(T) == 8; };
// strided tiles are N rows of 128 bytes: as N separate bulk copies they lose (Y pass 0.467 -> 0.82 ms at 308x616x616, measured);
// they keep 16-byte cp.async until the tile is ONE tensor-map copy. The contiguous Z lines (1.2-9.9 KB per copy) gain 2-4 %.
template <typename T> struct UseBulkStrided { static constexpr bool value = false; };

// Programmatic dependent launch (PDL): a kernel lets its successor in the stream start launching as soon as
// all of its own blocks are resident (launch_dependents), and waits for its predecessor to complete and flush
// before touching dependent data (wait). The successor's blocks fill the SM slots freed by the predecessor's
// finished blocks and run their prologue (twiddle tables) in the predecessor's tail. Both are no-ops when the
// kernel is launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    static const bool enabled = [] { const char* e = getenv("ADMP_PDL"); return !(e && atoi(e) == 0); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // measured on B200: +12 % on back-to-back stream launches, -3 % inside the captured SCF graph (the early
    // blocks of the successor hold SM slots while they wait) - so only outside stream capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    cfg.numAttrs = (enabled && cs == cudaStreamCaptureStatusNone) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// 1/x to ~1 ulp from the SFU approximation and two Newton steps (x normal, positive; x = 0 gives NaN:
// callers select around it). Replaces the ~20-instruction IEEE division in the per-point influence function.
__device__ __forceinline__ double fast_rcp(double x) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)x));
    double r = (double)rf;
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// resident blocks per SM the kernels are compiled for: caps the register count at 128 per thread for the
// two-stage transforms and 168 for the three-stage ones (their in-place middle stage keeps a butterfly's
// outputs in registers across a barrier)
template <int NT, int R3 = 1> struct MinBlocks {
    static constexpr int budget = R3 == 1 ? 512 : 384;      // threads per SM at 128 / 168 registers per thread
    static constexpr int NTr = (NT + 31) / 32 * 32;
    // three-stage transforms in 128-thread blocks (fewer threads per line, several butterflies per thread): two blocks per SM at
    // up to 255 registers - two independent barrier domains instead of one 224-thread block
    static constexpr int value = (R3 > 1 && NTr == 128 && NT == 128) ? 2 : ((budget / NTr) > 0 ? (budget / NTr) : 1);
};

// ------------------------------------------------------------------------------------------ one Stockham stage
// Butterflies jj = j, j + JT, ... < N/R of the stage with stride NS (= product of the earlier radices):
//   v[t] = in(jj + t*N/R) * W_N^(t * k * N/(NS*R)),  k = jj mod NS ;  DFT_R ;  out((jj-k)*R + k + t*NS)
// run() leaves the results in registers, put() hands them to `out`: the caller places a barrier in between
// when `out` overwrites the buffer `in` reads (in-place exchange).
// `tws`: compact twiddle table of the stage, entry (t-1)*NS + k = W_N^(t*k*N/(NS*R)) (unused when NS == 1).
template <typename T, int R, int SIGN, int N, int NS, int JT>
struct FStage {
    static constexpr int m = N / R, iters = (m + JT - 1) / JT;
    cx<T> v[iters][R];
    template <typename In>
    __device__ __forceinline__ void run(int j, const cx<T>* __restrict__ tws, In in) {
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int jj = j + it * JT;
            if ((it + 1) * JT <= m || jj < m) {
                const int k = (NS == 1) ? 0 : (jj % NS);
                v[it][0] = in(jj);
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    cx<T> x = in(jj + t * m);
                    if (NS > 1) {
                        cx<T> w = tws[(t - 1) * NS + k];
                        if (SIGN < 0) w.y = -w.y;
                        x = cmul(x, w);
                    }
                    v[it][t] = x;
                }
                Dft<T, R, SIGN>::run(v[it]);
            }
            if (it + 1 < iters) asm volatile("" ::: "memory");       // keep the next butterfly's loads behind this one (registers)
        }
    }
    // last forward stage only (NS == N/R, so butterfly jj holds the points jj + t*m): scale them and run the
    // first stage of an inverse transform of the same radix on them, in registers
    template <typename Scale>
    __device__ __forceinline__ void scale_inverse(int j, Scale f) {
        static_assert(NS * R == N && SIGN == 1, "scale_inverse: last forward stage only");
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int jj = j + it * JT;
            if ((it + 1) * JT <= m || jj < m) {
#pragma unroll
                for (int t = 0; t < R; ++t) v[it][t] = f(jj + t * m, v[it][t]);
                Dft<T, R, -1>::run(v[it]);
            }
        }
    }
    // outputs of a first stage (stride 1): butterfly jj -> positions jj*R + t
    template <typename Out>
    __device__ __forceinline__ void put_first(int j, Out out) const {
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int jj = j + it * JT;
            if ((it + 1) * JT <= m || jj < m) {
#pragma unroll
                for (int t = 0; t < R; ++t) out(jj * R + t, v[it][t]);
            }
        }
    }
    template <typename Out>
    __device__ __forceinline__ void put(int j, Out out) const {
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int jj = j + it * JT;
            if ((it + 1) * JT <= m || jj < m) {
                const int k = (NS == 1) ? 0 : (jj % NS);
                const int base = (jj - k) * R + k;
#pragma unroll
                for (int t = 0; t < R; ++t) out(base + t * NS, v[it][t]);
            }
        }
    }
};

// number of entries of the per-stage twiddle tables
template <int R1, int R2, int R3> struct TwGeom {
    static constexpr int N2 = (R2 - 1) * R1;                          // stage 2: NS = R1
    static constexpr int N3 = R3 > 1 ? (R3 - 1) * R1 * R2 : 0;        // stage 3: NS = R1*R2
    static constexpr int TOTAL = N2 + N3;
};
// fills tw2 / tw3 from the global table gtw[i] = exp(-2 pi i / (N*TWMUL))
template <typename T, int R1, int R2, int R3, int TWMUL>
__device__ __forceinline__ void build_twiddles(cx<T>* tw2, cx<T>* tw3, const cx<T>* __restrict__ gtw, int nthreads) {
    constexpr int N = R1 * R2 * R3;
    for (int i = threadIdx.x; i < TwGeom<R1, R2, R3>::N2; i += nthreads) {
        const int t = i / R1 + 1, k = i - (t - 1) * R1;
        tw2[i] = gtw[t * k * (N / (R1 * R2)) * TWMUL];
    }
    if (R3 > 1) {
        for (int i = threadIdx.x; i < TwGeom<R1, R2, R3>::N3; i += nthreads) {
            const int t = i / (R1 * R2) + 1, k = i - (t - 1) * (R1 * R2);
            tw3[i] = gtw[t * k * TWMUL];
        }
    }
}

// remaining stages (2, and 3 when R3 > 1) of a line FFT whose stage-1 output sits in buffer A:
// the middle stage of a three-stage transform exchanges in place in A, the last stage hands its outputs to
// `last`; LAST_INPLACE (last writes A again) puts a barrier between the last stage's reads and writes.
template <typename T, int R1, int R2, int R3, int SIGN, int JT, bool LAST_INPLACE, typename LdA, typename StA, typename Last>
__device__ __forceinline__ void fft_tail(int j, bool live, const cx<T>* __restrict__ tw2, const cx<T>* __restrict__ tw3, LdA ldA, StA stA,
                                         Last last) {
    constexpr int N = R1 * R2 * R3;
    if (R3 == 1) {
        FStage<T, R2, SIGN, N, R1, JT> s;
        if (live) s.run(j, tw2, ldA);
        if (LAST_INPLACE) __syncthreads();
        if (live) s.put(j, last);
    } else {
        {
            FStage<T, R2, SIGN, N, R1, JT> s;
            if (live) s.run(j, tw2, ldA);
            __syncthreads();
            if (live) s.put(j, stA);
            __syncthreads();
        }
        constexpr int R3e = R3 > 1 ? R3 : 2;
        FStage<T, R3e, SIGN, N, R1 * R2, JT> s;
        if (live) s.run(j, tw3, ldA);
        if (LAST_INPLACE) __syncthreads();
        if (live) s.put(j, last);
    }
}
// stage 1: src -> dst (different buffers; no barrier inside)
template <typename T, int R1, int R2, int R3, int SIGN, int JT, typename Ld, typename St>
__device__ __forceinline__ void fft_head(int j, bool live, Ld ld, St st) {
    if (live) {
        FStage<T, R1, SIGN, R1 * R2 * R3, 1, JT> s;
        s.run(j, nullptr, ld);
        s.put(j, st);
    }
}
// ------------------------------------------------------------------------------------------ prime-factor (Good-Thomas) stages
// All sizes of the mesh family factor into PAIRWISE COPRIME radices (77 = 11*7, 154 = 11*14, 308 = 11*7*4, 616 = 11*7*8,
// 1232 = 11*7*16), so the line FFT of the strided passes is a pure R1 x R2 x R3 multi-dimensional DFT - no twiddle factors:
//   input   n = (n1 S1 + n2 S2 + n3 S3) mod N,  S_d = N / R_d                       (Good's map, natural order in)
//   output  k = (k1 T1 + k2 T2 + k3 T3) mod N,  T_d = S_d * (S_d^-1 mod R_d)         (CRT map, natural order out)
//   W_N^(n k) = prod_d W_Rd^(n_d k_d)           (cross terms S_d T_e are multiples of N, S_d T_d = 1 mod R_d)
// The work buffer holds the multi-index array row-major, w(i1, i2, i3) at i1 + R1 (i2 + R2 i3). The first dimension gathers
// its inputs from the natural-order tile through Good's map, the middle dimension works IN PLACE (every butterfly reads and
// writes the same positions: no barrier between its loads and stores), the last dimension scatters through the CRT map. In
// the fused X pass the frequency side never needs natural order: the influence tables are permuted once per block, and the
// inverse runs the dimensions backwards and leaves through Good's map again. Versus the Stockham stages above: no twiddle
// tables / complex multiplies (16 % of the FP64 instructions of a 308-point X pass), 5 instead of 9 barriers per X tile.
#ifndef ADMP_FFT_PFA
#define ADMP_FFT_PFA 1
#endif
constexpr int pfa_gcd(int a, int b) { return b == 0 ? a : pfa_gcd(b, a % b); }
constexpr int pfa_inv(int a, int m) {
    int r = 0;
    for (int x = 1; x < m; ++x)
        if ((a % m) * x % m == 1) r = x;
    return r;
}
template <int R1, int R2, int R3> struct Pfa {
    static constexpr int N = R1 * R2 * R3;
    static constexpr bool value = ADMP_FFT_PFA && pfa_gcd(R1, R2) == 1 && pfa_gcd(R1, R3) == 1 && pfa_gcd(R2, R3) == 1;
    // stand-alone Y passes: measured on B200 against the Stockham stages - 154 points 18.3 -> 17.2 us, 616 points 0.444 -> 0.456 ms
    // (HBM-bound either way), so only the two-factor sizes use it there; the fused X pass gains at every size (308: 0.919 -> 0.747 ms)
#ifndef ADMP_Y_PFA3
#define ADMP_Y_PFA3 1
#endif
    static constexpr bool strided = value && (R3 == 1 || ADMP_Y_PFA3);
    static constexpr int S1 = N / R1, S2 = N / R2, S3 = N / R3;
    static constexpr int T1 = (S1 * pfa_inv(S1, R1)) % N, T2 = (S2 * pfa_inv(S2, R2)) % N, T3 = R3 > 1 ? (S3 * pfa_inv(S3, R3)) % N : 0;
    // butterflies of dimension 1: b = i2 + R2 i3
    static __device__ __forceinline__ int rm1(int b) { return R1 * b; }
    static __device__ __forceinline__ int rm1_off(int bs, int t) { return bs + t; }
    static __device__ __forceinline__ int good1(int b) {
        const int i3 = b / R2, i2 = b - i3 * R2;
        int x = i2 * S2 + (R3 > 1 ? i3 * S3 : 0);
        return x >= N ? x - N : x;
    }
    static __device__ __forceinline__ int good1_off(int bs, int t) {
        const int p = bs + t * S1;
        return p >= N ? p - N : p;
    }
    // the same dimension for line-major tiles (Z passes: lanes = consecutive butterflies of ONE line): b = i3 + R3 i2, so that
    // consecutive lanes are S3 = R1 R2 (odd) points apart on both sides - no shared-memory bank conflicts
    static __device__ __forceinline__ int rm1z(int b) {
        if (R3 == 1) return R1 * b;
        const int i2 = b / R3, i3 = b - i2 * R3;
        return R1 * (i2 + R2 * i3);
    }
    static __device__ __forceinline__ int good1z(int b) {
        if (R3 == 1) return b * S2;
        const int i2 = b / R3, i3 = b - i2 * R3;
        const int x = i2 * S2 + i3 * S3;
        return x >= N ? x - N : x;
    }
    // row-major position of frequency k (k = k_d mod R_d)
    static __device__ __forceinline__ int pos_of_freq(int k) { return k % R1 + R1 * (k % R2 + R2 * (R3 > 1 ? k % R3 : 0)); }
    // butterflies of dimension 2 (three factors: middle dimension): b = i1 + R1 i3
    static __device__ __forceinline__ int rm2(int b) {
        const int i3 = b / R1;
        return b + (R2 - 1) * R1 * i3;
    }
    static __device__ __forceinline__ int rm2_off(int bs, int t) { return bs + t * R1; }
    // last dimension (3 when R3 > 1, else 2): b = row-major index of the other dimensions, stride = their product
    static constexpr int RL = R3 > 1 ? R3 : R2, SL = N / RL, TLAST = R3 > 1 ? T3 : T2;
    static __device__ __forceinline__ int rml(int b) { return b; }
    static __device__ __forceinline__ int rml_off(int bs, int t) { return bs + t * SL; }
    static __device__ __forceinline__ int crtl(int b) {
        if (R3 > 1) {
            const int i2 = b / R1, i1 = b - i2 * R1;
            return (i1 * T1 + i2 * T2) % N;
        }
        return (b * T1) % N;
    }
    static __device__ __forceinline__ int crtl_off(int bs, int t) {
        const int p = bs + (t * TLAST) % N;
        return p >= N ? p - N : p;
    }
    // frequency of row-major position p (set-up of the permuted influence tables)
    static __device__ __forceinline__ int freq(int p) {
        const int i1 = p % R1, r = p / R1, i2 = r % R2, i3 = r / R2;
        return (i1 * T1 + i2 * T2 + i3 * T3) % N;
    }
};

// one dimension of the multi-dimensional DFT: butterflies b = j, j + JT, ... < NB of radix R; `base(b)` / `off(base, t)` give the
// position of point t of butterfly b behind `in` / `out`. Every butterfly is loaded, transformed and stored on its own (in place or
// into another buffer: no butterfly reads what another one writes within a dimension), so a thread holds ONE butterfly in registers
// however many it processes.
template <typename T, int R, int SIGN, int NB, int JT>
struct PStage {
    static constexpr int iters = (NB + JT - 1) / JT;
    template <typename BaseI, typename OffI, typename In, typename BaseO, typename OffO, typename Out>
    static __device__ __forceinline__ void run(int j, BaseI basei, OffI offi, In in, BaseO baseo, OffO offo, Out out) {
#pragma unroll 1
        for (int b = j; b < NB; b += JT) {                           // rolled: one butterfly's registers and code, whatever NB / JT
            cx<T> v[R];
            const int bi = basei(b);
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = in(offi(bi, t));
            Dft<T, R, SIGN>::run(v);
            const int bo = baseo(b);
#pragma unroll
            for (int t = 0; t < R; ++t) out(offo(bo, t), v[t]);
        }
    }
    // forward DFT, v <- f(position, v), inverse DFT, back to the same positions (frequency-side scaling between a forward and an
    // inverse transform, in registers)
    template <typename Base, typename Off, typename In, typename Scale, typename Out>
    static __device__ __forceinline__ void run_scaled(int j, Base base, Off off, In in, Scale f, Out out) {
#pragma unroll 1
        for (int b = j; b < NB; b += JT) {
            cx<T> v[R];
            const int bs = base(b);
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = in(off(bs, t));
            Dft<T, R, SIGN>::run(v);
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = f(off(bs, t), v[t]);
            Dft<T, R, -SIGN>::run(v);
#pragma unroll
            for (int t = 0; t < R; ++t) out(off(bs, t), v[t]);
        }
    }
};

// ------------------------------------------------------------------------------------------ strided passes (Y, X)
template <typename T, int N, int TL, int JT>
__device__ __forceinline__ void issue_tile(const StrideGeom& g, int tile, const cx<T>* __restrict__ spec, cx<T>* dst, int l, int j) {
    const int o = tile / g.tiles, t = tile - o * g.tiles;
    const int c0 = t * TL;
    const int nl = min(TL, g.n_inner - c0);
    const cx<T>* base = spec + (size_t)o * g.outer_stride + c0 + l;
    if (l < nl) {
#pragma unroll 4
        for (int pos = j; pos < N; pos += JT) cp_async<sizeof(cx<T>)>(dst + pos * TL + l, base + (size_t)pos * g.line_stride);
    }
    cp_async_commit();
}

// the tile as N / BOXN tensor-map copies issued by ONE thread: the TMA unit walks the rows (128 bytes each, any stride), the
// other threads spend no instruction and no register on it; columns beyond the plane edge are zero-filled by the unit and the
// full box is always accounted on the mbarrier
template <typename T, int N, int TL>
__device__ __forceinline__ void issue_tile_tma(const StrideGeom& g, int tile, const CUtensorMap* tm, cx<T>* dst, uint64_t* bar) {
    constexpr int BOXN = tma_box_rows(N, TL) > 0 ? tma_box_rows(N, TL) : N;
    if (threadIdx.x == 0) {
        const int o = tile / g.tiles, t = tile - o * g.tiles;
        fence_async_smem();
        mbar_expect_tx(bar, (unsigned)(N * TL * sizeof(cx<T>)));
#pragma unroll
        for (int b = 0; b < N / BOXN; ++b) tma_load_3d(dst + b * BOXN * TL, tm, 2 * t * TL, b * BOXN, o, bar);
    }
}

// the same tile through the TMA unit: one bulk copy per position (nl consecutive lines = one contiguous run)
template <typename T, int N, int TL>
__device__ __forceinline__ void issue_tile_bulk(const StrideGeom& g, int tile, const cx<T>* __restrict__ spec, cx<T>* dst, int nthreads,
                                                uint64_t* bar) {
    const int o = tile / g.tiles, t = tile - o * g.tiles;
    const int c0 = t * TL;
    const int nl = min(TL, g.n_inner - c0);
    const cx<T>* base = spec + (size_t)o * g.outer_stride + c0;
    const unsigned bytes = (unsigned)(nl * sizeof(cx<T>));
    fence_async_smem();
    if (threadIdx.x == 0) mbar_expect_tx(bar, bytes * N);
    for (int pos = threadIdx.x; pos < N; pos += nthreads) bulk_g2s(dst + pos * TL, base + (size_t)pos * g.line_stride, bytes, bar);
}

// X pass over an x-slab decomposed spectrum (PeerTab): point `pos` of every line lives in the buffer of rank
// owner(pos); `sbase[pos]` (shared memory) holds that buffer's base pointer, element offsets are unchanged.
template <typename T, int N, int TL, int JT>
__device__ __forceinline__ void issue_tile_peer(const StrideGeom& g, int tile, cx<T>* const* sbase, cx<T>* dst, int l, int j) {
    const int c0 = tile * TL;
    const int nl = min(TL, g.n_inner - c0);
    if (l < nl) {
#pragma unroll 4
        for (int pos = j; pos < N; pos += JT) cp_async<sizeof(cx<T>)>(dst + pos * TL + l, sbase[pos] + (size_t)pos * g.line_stride + c0 + l);
    }
    cp_async_commit();
}

// resident blocks of the prime-factor passes: the stages hold one butterfly per thread (~150 registers at radix 11), so two
// blocks of a three-stage transform fit (shared memory - two tiles per block - allows no more at 8 lines of 308 points)
#ifndef ADMP_X_MINBLOCKS
#define ADMP_X_MINBLOCKS 2
#endif
template <int NT, int R1, int R2, int R3> struct XMinBlocks {
    static constexpr int base = MinBlocks<NT, R3>::value;
    static constexpr int value = (Pfa<R1, R2, R3>::value && R3 > 1 && NT <= 256 && base < ADMP_X_MINBLOCKS) ? ADMP_X_MINBLOCKS : base;
};
template <typename T, int R1, int R2, int R3, int SIGN, int TL, int JT, bool TMA>
__global__ void __launch_bounds__(TL* JT, (Pfa<R1, R2, R3>::strided ? XMinBlocks<TL * JT, R1, R2, R3>::value : MinBlocks<TL * JT, R3>::value))
fast_strided_kernel(StrideGeom g, int ntiles, cx<T>* __restrict__ spec, const cx<T>* __restrict__ gtw, const __grid_constant__ CUtensorMap tmap) {
    constexpr int N = R1 * R2 * R3, TILE = N * TL, NT = TL * JT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* I = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* A = I + TILE;
    cx<T>* tw2 = A + TILE;
    cx<T>* tw3 = tw2 + TwGeom<R1, R2, R3>::N2;
    constexpr bool BULK = TMA || UseBulkStrided<T>::value;
    __shared__ uint64_t bar;
    unsigned parity = 0;
    const int l = threadIdx.x % TL, j = threadIdx.x / TL;
    int tile = blockIdx.x;
    pdl_launch_dependents();
    if (BULK && threadIdx.x == 0) mbar_init(&bar, 1);
    if (!Pfa<R1, R2, R3>::strided) build_twiddles<T, R1, R2, R3, 1>(tw2, tw3, gtw, NT);
    if (BULK) __syncthreads();
    pdl_wait();
    auto issue = [&](int t) {
        if (TMA) issue_tile_tma<T, N, TL>(g, t, &tmap, I, &bar);
        else if (BULK) issue_tile_bulk<T, N, TL>(g, t, spec, I, NT, &bar);
        else issue_tile<T, N, TL, JT>(g, t, spec, I, l, j);
    };
    if (tile < ntiles) issue(tile);
    cx<T>* a = A + l;
    cx<T>* c = I + l;
    for (; tile < ntiles; tile += gridDim.x) {
        if (BULK) { mbar_wait(&bar, parity); parity ^= 1; }
        else cp_async_wait_all();
        __syncthreads();                       // tile landed in I; A free (previous tile's last stage has read it)
        const int o = tile / g.tiles, t = tile - o * g.tiles;
        const int c0 = t * TL;
        const bool live = l < g.n_inner - c0;
        cx<T>* out = spec + (size_t)o * g.outer_stride + c0 + l;
        const size_t ls = g.line_stride;
        if (Pfa<R1, R2, R3>::strided) {
            using P = Pfa<R1, R2, R3>;
            auto ldI = [&](int pos) { return c[pos * TL]; };
            auto ldA = [&](int pos) { return a[pos * TL]; };
            auto stA = [&](int pos, cx<T> v) { a[pos * TL] = v; };
            // dimension 1: natural-order tile (Good's map) -> row-major work buffer
            if (live) PStage<T, R1, SIGN, N / R1, JT>::run(j, P::good1, P::good1_off, ldI, P::rm1, P::rm1_off, stA);
            if (BULK) fence_async_smem();
            __syncthreads();                   // I consumed
            if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
            if (R3 > 1) {                      // middle dimension, in place (no barrier between its loads and stores)
                if (live) PStage<T, R2, SIGN, N / R2, JT>::run(j, P::rm2, P::rm2_off, ldA, P::rm2, P::rm2_off, stA);
                __syncthreads();
            }
            // last dimension: results leave through the CRT map, in natural order
            if (live)
                PStage<T, P::RL, SIGN, N / P::RL, JT>::run(j, P::rml, P::rml_off, ldA, P::crtl, P::crtl_off,
                                                           [&](int pos, cx<T> v) { out[(size_t)pos * ls] = v; });
            continue;
        }
        fft_head<T, R1, R2, R3, SIGN, JT>(j, live, [&](int pos) { return c[pos * TL]; }, [&](int pos, cx<T> v) { a[pos * TL] = v; });
        if (BULK) fence_async_smem();
        __syncthreads();                       // I consumed
        if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
        fft_tail<T, R1, R2, R3, SIGN, JT, false>(j, live, tw2, tw3, [&](int pos) { return a[pos * TL]; },
                                                 [&](int pos, cx<T> v) { a[pos * TL] = v; },
                                                 [&](int pos, cx<T> v) { out[(size_t)pos * ls] = v; });
    }
}

// general influence function (any kind, any cell, optional virial sums): cold path of the fused X pass
__device__ __noinline__ double influence_general(const BoxInfo* Bp, const ConvTables* tbp, int kind, double kap, int i1, int i2, int i3,
                                                 double s2, int want_vir, double* acc_t) {
    const BoxInfo& B = *Bp;
    const ConvTables& tb = *tbp;
    const bool ortho = *tb.ortho != 0;
    const bool single = (i3 == 0) || (2 * i3 == B.K[2]);
    if (want_vir) {
        const Influence f = influence<true>(B, tb, ortho, kap, kind, i1, i2, i3);
        double a[6] = {0, 0, 0, 0, 0, 0};
        virial_terms(B, f.kv, i1, i2, i3, single, f.dg * s2, a);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc_t[k] += a[k];
        return f.g;
    }
    return influence<false>(B, tb, ortho, kap, kind, i1, i2, i3).g;
}

// X-forward, multiply by 2*scale*C_k/theta_k^2 with energy (+ virial) accumulation, X-inverse
// (admp/recip.py:410-426 fused with its adjoint). QUICK = Coulomb energy only (every SCF cycle): separable
// tables of conv_tables_kernel + one reciprocal per point when the cell is orthorhombic (device flag),
// inline exp otherwise; !QUICK = any kind / virial sums through influence_general.
// PEER: the spectrum is x-slab decomposed over the GPUs of the NVLink domain; this rank transforms the lines
// of tiles [tile0, ntiles) by loading / storing every point from / to the rank that owns its x plane (cp.async
// and stores on peer-mapped memory): the all-to-all transposes of a slab FFT are fused into the X pass.
// per-line virial sums -> the six T_ab accumulators (xx, xy, xz, yy, yz, zz)
__device__ __forceinline__ void fold_virial(double (&acc_t)[6], double Sxx, double Sxy, double Sxz, double Sb, double ky, double kz,
                                            double pw, double p1) {
    acc_t[0] += (1.0 + pw) * Sxx;
    acc_t[1] += Sxy;
    acc_t[2] += Sxz;
    acc_t[3] += (1.0 + pw) * ky * ky * Sb;
    acc_t[4] += (1.0 - pw * p1) * ky * kz * Sb;
    acc_t[5] += (1.0 + pw) * kz * kz * Sb;
}
// VIR (QUICK only): the Coulomb pass that also accumulates the k-space virial sums T_ab = sum_k w dC/dk^2 |S|^2 / theta^2 k_a k_b
// (the one pass after the SCF loop), from the same separable tables on an orthorhombic cell: ~11 more FP64 operations per point
// instead of the general influence function (exp, divisions, a call per point: 3x the time of the SCF-cycle pass).
template <typename T, int R1, int R2, int R3, int TL, int JT, bool QUICK, bool PEER, bool TMA = false, bool VIR = false>
__global__ void __launch_bounds__(TL* JT, XMinBlocks<TL * JT, R1, R2, R3>::value)
fast_x_conv_kernel(StrideGeom g, int tile0, int ntiles, const BoxInfo* __restrict__ Bp, T kappa, int kind, ConvTables tb,
                   cx<T>* __restrict__ spec, const cx<T>* __restrict__ gtw, double* __restrict__ scalars, int want_vir, PeerTab peers,
                   int local_reads, const __grid_constant__ CUtensorMap tmap) {
    constexpr int N = R1 * R2 * R3, TILE = N * TL, NT = TL * JT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[7 * ((NT + 31) / 32)];
    cx<T>* I = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* A = I + TILE;
    cx<T>* tw2 = A + TILE;
    cx<T>* tw3 = tw2 + TwGeom<R1, R2, R3>::N2;
    // the inverse runs with the radices in reverse order (see the loop), so it needs its own twiddle tables
    constexpr int Q1 = R3 > 1 ? R3 : R2, Q2 = R3 > 1 ? R2 : R1, Q3 = R3 > 1 ? R1 : 1;
    cx<T>* itw2 = tw3 + TwGeom<R1, R2, R3>::N3;
    cx<T>* itw3 = itw2 + TwGeom<Q1, Q2, Q3>::N2;
    double* sek = reinterpret_cast<double*>(itw3 + TwGeom<Q1, Q2, Q3>::N3);    // exp(-k1^2/4kappa^2)/theta_1^2  (ortho) | 1/theta_1^2
    double* sk2 = sek + N;                                                    // k1^2 (ortho) | signed index m1
    cx<T>** sbase = reinterpret_cast<cx<T>**>(sk2 + N);                       // PEER only: owner buffer of x plane pos
    double* skx = sk2 + N;                                                    // VIR only (never with PEER): signed k1, Nyquist parity p0
    double* sp0 = skx + N;
    const BoxInfo& B = *Bp;
    constexpr bool BULK = TMA || UseBulkStrided<T>::value;
    __shared__ uint64_t bar;
    unsigned parity = 0;
    const bool bulk = BULK && !(PEER && !local_reads);      // peer-memory loads keep cp.async
    const int l = threadIdx.x % TL, j = threadIdx.x / TL;
    int tile = tile0 + blockIdx.x;
    pdl_launch_dependents();
    if (BULK && threadIdx.x == 0) mbar_init(&bar, 1);
    constexpr bool PFA = Pfa<R1, R2, R3>::value;
    using P = Pfa<R1, R2, R3>;
    int* sfreq = reinterpret_cast<int*>(tw2);                 // PFA: frequency of row-major position p (the twiddle tables are unused)
    if (!PFA) {
        build_twiddles<T, R1, R2, R3, 1>(tw2, tw3, gtw, NT);
        build_twiddles<T, Q1, Q2, Q3, 1>(itw2, itw3, gtw, NT);
    } else {
        for (int i = threadIdx.x; i < N; i += NT) sfreq[i] = P::freq(i);
    }
    if (BULK) __syncthreads();
    if (PEER) {
        for (int i = threadIdx.x; i < N; i += NT) sbase[i] = reinterpret_cast<cx<T>*>(peers.base[peers.owner(i)]);
        __syncthreads();
    }
    pdl_wait();
    // local_reads: the peers' planes of this rank's columns were pulled into `spec` beforehand (bulk copies over
    // NVLink, pipelined by the host side): loads are local, only the stores go to the owners
    auto issue = [&](int t) {
        if (PEER && !local_reads) issue_tile_peer<T, N, TL, JT>(g, t, sbase, I, l, j);
        else if (TMA) issue_tile_tma<T, N, TL>(g, t, &tmap, I, &bar);
        else if (bulk) issue_tile_bulk<T, N, TL>(g, t, spec, I, NT, &bar);
        else issue_tile<T, N, TL, JT>(g, t, spec, I, l, j);
    };
    if (tile < ntiles) issue(tile);
    const bool ortho = (*tb.ortho != 0);
    if (QUICK) {                                              // PFA: tables in the order of the work buffer (row-major multi-index)
        for (int i = threadIdx.x; i < N; i += NT) {
            const int f = PFA ? P::freq(i) : i;
            sek[i] = ortho ? tb.ek[0][f] : tb.bt[0][f];
            sk2[i] = ortho ? tb.k2[0][f] : (double)kint(f, N);
            if (VIR) {
                skx[i] = 6.283185307179586 * (double)kint(f, N) * B.inv[0];
                sp0[i] = (2 * f == N) ? 1.0 : -1.0;
            }
        }
    }
    const int K2 = B.K[1], K3 = B.K[2], K3h = K3 / 2 + 1;
    const double scale = (kind == ADMP_CK_COULOMB) ? ADMP_DIEL : 1.0;
    const double kap = (double)kappa;
    const double pref = 2.0 * scale * 6.283185307179586 / B.vol;
    const double q4k = -1.0 / (4.0 * kap * kap);
    double acc_e = 0.0, acc_t[6] = {0, 0, 0, 0, 0, 0};
    cx<T>* a = A + l;
    cx<T>* c = I + l;
    // Per tile: stage 1 moves I -> A and the next tile is fetched into I; everything else happens in A.
    // The forward transform's last stage leaves thread jj with the points {jj + t*N/R}: exactly the inputs of a
    // first inverse stage of the same radix, so the scaling and that inverse stage run in registers (no
    // exchange through shared memory, no barrier) and the inverse continues with the radices in reverse order.
    for (; tile < ntiles; tile += gridDim.x) {
        if (bulk) { mbar_wait(&bar, parity); parity ^= 1; }
        else cp_async_wait_all();
        __syncthreads();
        const int c0 = tile * TL;
        const bool live = l < g.n_inner - c0;
        const int cl = live ? c0 + l : 0;
        const int i2 = cl / K3h, i3 = cl - i2 * K3h;
        const double wgt = ((i3 == 0) || (2 * i3 == K3)) ? 1.0 : 2.0;
        const bool origin_line = (i2 == 0 && i3 == 0);
        // per-line constants of the Coulomb influence function
        double e23 = 0.0, k23 = 1.0, kb[3] = {0, 0, 0};
        if (QUICK) {
            if (ortho) {
                e23 = pref * tb.ek[1][i2] * tb.ek[2][i3];         // 2*scale*(2 pi/V) * the y, z factors
                k23 = tb.k2[1][i2] + tb.k2[2][i3];
            } else {
                e23 = pref * tb.bt[1][i2] * tb.bt[2][i3];
                const double m2 = kint(i2, K2), m3 = i3;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) kb[cc] = 6.283185307179586 * (m2 * B.inv[3 + cc] + m3 * B.inv[6 + cc]);
            }
        }
        double acc_line = 0.0;
        // VIR: per-line factors of the virial products (influence.cuh virial_terms on an orthorhombic cell: k = (kx, ky, kz), Hermitian
        // partner (p0 kx, p1 ky, -kz) for weight-2 points) and the per-line sums
        double vky = 0.0, vkz = 0.0, vxyB = 0.0, vxzB = 0.0, Sxx = 0.0, Sxy = 0.0, Sxz = 0.0, Sb = 0.0;
        const double pw = (wgt == 2.0) ? 1.0 : 0.0;
        const double p1 = (2 * i2 == K2) ? 1.0 : -1.0;
        if (VIR) {
            vky = 6.283185307179586 * (double)kint(i2, K2) * B.inv[4];
            vkz = 6.283185307179586 * (double)i3 * B.inv[8];
            vxyB = vky * pw * p1;
            vxzB = -vkz * pw;
        }
        cx<T>* out = spec + c0 + l;
        const size_t ls = g.line_stride;
        auto scale_point = [&](int i1, cx<T> s) -> cx<T> {
            const double s2 = (double)s.x * s.x + (double)s.y * s.y;
            double gg;                                        // 2*scale*C_k/theta_k^2
            if (QUICK) {
                if (ortho) {
                    const double rk = fast_rcp(sk2[i1] + k23);
                    gg = e23 * sek[i1] * rk;
                    if (VIR && !(origin_line && i1 == 0)) {
                        // b = scale * dC/dk^2 / theta^2 * |S|^2 = -gg/2 (1/k^2 + 1/4kappa^2) |S|^2
                        const double b = -0.5 * gg * (rk - q4k) * s2, bx = b * skx[i1];
                        Sxx = fma(bx, skx[i1], Sxx);
                        Sxy = fma(bx, fma(vxyB, sp0[i1], vky), Sxy);
                        Sxz = fma(bx, fma(vxzB, sp0[i1], vkz), Sxz);
                        Sb += b;
                    }
                } else if (VIR) {
                    gg = 2.0 * scale * influence_general(Bp, &tb, kind, kap, PFA ? sfreq[i1] : i1, i2, i3, s2 * scale, want_vir, acc_t);
                } else {
                    const double m1 = sk2[i1];
                    const double kx = kb[0] + 6.283185307179586 * m1 * B.inv[0], ky = kb[1] + 6.283185307179586 * m1 * B.inv[1],
                                 kz = kb[2] + 6.283185307179586 * m1 * B.inv[2];
                    const double ksq = kx * kx + ky * ky + kz * kz;
                    gg = e23 * sek[i1] * exp(ksq * q4k) * fast_rcp(ksq);
                }
                if (origin_line && i1 == 0) gg = 0.0;         // gamma point dropped (recip.py:416)
            } else {
                gg = 2.0 * scale * influence_general(Bp, &tb, kind, kap, PFA ? sfreq[i1] : i1, i2, i3, s2 * scale, want_vir, acc_t);
            }
            acc_line = fma(gg, s2, acc_line);
            const T gt = (T)gg;
            return {s.x * gt, s.y * gt};
        };
        auto ldA = [&](int pos) { return a[pos * TL]; };
        auto stA = [&](int pos, cx<T> v) { a[pos * TL] = v; };
        if (PFA) {
            // forward: dimension 1 (natural tile -> row-major), [2 in place], last dimension + scaling + its inverse in
            // registers; inverse: [2 in place], dimension 1 leaves through Good's map in natural order. `scale_point` sees
            // row-major positions: its tables were permuted at set-up (position 0 is frequency 0).
            if (live) PStage<T, R1, 1, N / R1, JT>::run(j, P::good1, P::good1_off, [&](int pos) { return c[pos * TL]; }, P::rm1, P::rm1_off, stA);
            if (bulk) fence_async_smem();
            __syncthreads();                   // I consumed: fetch the next tile while this one is transformed
            if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
            if (R3 > 1) {
                if (live) PStage<T, R2, 1, N / R2, JT>::run(j, P::rm2, P::rm2_off, ldA, P::rm2, P::rm2_off, stA);
                __syncthreads();
            }
            if (live) PStage<T, P::RL, 1, N / P::RL, JT>::run_scaled(j, P::rml, P::rml_off, ldA, scale_point, stA);
            __syncthreads();
            acc_e = fma(0.5 * wgt, acc_line, acc_e);
            if (VIR) fold_virial(acc_t, Sxx, Sxy, Sxz, Sb, vky, vkz, pw, p1);
            if (R3 > 1) {
                if (live) PStage<T, R2, -1, N / R2, JT>::run(j, P::rm2, P::rm2_off, ldA, P::rm2, P::rm2_off, stA);
                __syncthreads();
            }
            if (live) {
                if (PEER) {
                    const size_t col = (size_t)c0 + l;
                    PStage<T, R1, -1, N / R1, JT>::run(j, P::rm1, P::rm1_off, ldA, P::good1, P::good1_off,
                                                       [&](int pos, cx<T> v) { sbase[pos][(size_t)pos * ls + col] = v; });
                } else {
                    PStage<T, R1, -1, N / R1, JT>::run(j, P::rm1, P::rm1_off, ldA, P::good1, P::good1_off,
                                                       [&](int pos, cx<T> v) { out[(size_t)pos * ls] = v; });
                }
            }
            continue;
        }
        fft_head<T, R1, R2, R3, 1, JT>(j, live, [&](int pos) { return c[pos * TL]; }, stA);
        if (bulk) fence_async_smem();
        __syncthreads();                       // I consumed: fetch the next tile while this one is transformed
        if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
        if (R3 > 1) {                          // middle forward stage, in place in A
            FStage<T, R2, 1, N, R1, JT> s;
            if (live) s.run(j, tw2, ldA);
            __syncthreads();
            if (live) s.put(j, stA);
            __syncthreads();
        }
        {
            // last forward stage (radix Q1, stride N/Q1) + scaling + first inverse stage (radix Q1), in registers
            FStage<T, Q1, 1, N, N / Q1, JT> s;
            if (live) {
                s.run(j, R3 > 1 ? tw3 : tw2, ldA);
                s.scale_inverse(j, scale_point);
            }
            __syncthreads();
            if (live) s.put_first(j, stA);
            __syncthreads();
        }
        acc_e = fma(0.5 * wgt, acc_line, acc_e);              // E = scale * sum wgt g |S|^2 = sum wgt/2 * gg |S|^2
        if (VIR) fold_virial(acc_t, Sxx, Sxy, Sxz, Sb, vky, vkz, pw, p1);
        if (PEER) {
            const size_t col = (size_t)c0 + l;
            fft_tail<T, Q1, Q2, Q3, -1, JT, false>(j, live, itw2, itw3, ldA, stA,
                                                   [&](int pos, cx<T> v) { sbase[pos][(size_t)pos * ls + col] = v; });
        } else {
            fft_tail<T, Q1, Q2, Q3, -1, JT, false>(j, live, itw2, itw3, ldA, stA, [&](int pos, cx<T> v) { out[(size_t)pos * ls] = v; });
        }
    }
    double e1[1] = {acc_e};
    block_accumulate<1>(e1, red, scalars + ADMP_S_E_RECIP);
    if ((!QUICK || VIR) && want_vir) block_accumulate<6>(acc_t, red, scalars + ADMP_S_TK);
}

// ------------------------------------------------------------------------------------------ Z passes (contiguous lines)
// thread = (butterfly fastest, line); line-major tile [TL][LS] in shared memory, M = R1*R2*R3 = K3/2.
template <int M> struct ZGeom {
    static constexpr int LS = (M + 1) | 1;        // holds the M+1 half-spectrum points of a line; odd stride
};

// forward: K3 reals per line -> K3/2+1 complex (packed real FFT, tools/fft_model.py r2c)
template <typename T, int R1, int R2, int R3, int TL, int JT>
__global__ void __launch_bounds__(TL* JT, XMinBlocks<TL * JT, R1, R2, R3>::value)
fast_z_fwd_kernel(int nlines, int ntiles, const T* __restrict__ mesh, cx<T>* __restrict__ spec, const cx<T>* __restrict__ gtw, int zlc) {
    // zlc: 16-byte units per REAL line (M: the mesh has its own buffer; M + 1: the mesh lives in the spectrum buffer, line by line in
    // place - a tile is complete in shared memory before its spectrum lines are written, and no other block touches those lines)
    constexpr int M = R1 * R2 * R3, K3h = M + 1, LS = ZGeom<M>::LS, TILE = TL * LS, NT = TL * JT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* I = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* A = I + TILE;
    cx<T>* tw2 = A + TILE;
    cx<T>* tw3 = tw2 + TwGeom<R1, R2, R3>::N2;
    cx<T>* zt = tw3 + TwGeom<R1, R2, R3>::N3;       // M+1 entries: exp(-2 pi i k / K3)
    // Tiles whose TL lines fill a 128-byte bank row (8 lines of float64): thread = (line fastest, butterfly) - the lanes of a quarter
    // warp work on the SAME butterfly of 8 different lines, LS (odd) apart in shared memory: 8 distinct 16-byte bank groups whatever the
    // position pattern of the stage (Good gather, row-major, Stockham strides). With the butterfly index fastest this pass spent 50 %
    // of its stall samples on shared memory (ncu r2k: short scoreboard 34 %, MIO 16 %). Narrower tiles keep the butterfly index
    // fastest (line-fastest halves of two butterflies collide: measured 0.494 -> 0.553 ms with 4-line tiles at 308x616x616).
    constexpr bool LINE_FASTEST = TL * sizeof(cx<T>) >= 128;
    const int j = LINE_FASTEST ? threadIdx.x / TL : threadIdx.x % JT, l = LINE_FASTEST ? threadIdx.x % TL : threadIdx.x / JT;
    constexpr bool BULK = UseBulk<T>::value;
    __shared__ uint64_t bar;
    unsigned parity = 0;
    auto issue = [&](int tile) {
        const int L0 = tile * TL;
        const int nl = min(TL, nlines - L0);
        const cx<T>* src = reinterpret_cast<const cx<T>*>(mesh) + (size_t)L0 * zlc;
        if (BULK) {                                   // one bulk copy per line: K3 reals = M 16-byte units
            fence_async_smem();
            if (threadIdx.x == 0) mbar_expect_tx(&bar, (unsigned)(nl * M * sizeof(cx<T>)));
            if (threadIdx.x < nl) bulk_g2s(I + threadIdx.x * LS, src + (size_t)threadIdx.x * zlc, (unsigned)(M * sizeof(cx<T>)), &bar);
        } else {
            for (int e = threadIdx.x; e < nl * M; e += NT) {
                const int ll = e / M, pos = e - ll * M;
                cp_async<sizeof(cx<T>)>(I + ll * LS + pos, src + (size_t)ll * zlc + pos);
            }
            cp_async_commit();
        }
    };
    int tile = blockIdx.x;
    pdl_launch_dependents();
    if (BULK && threadIdx.x == 0) mbar_init(&bar, 1);
    // (three-factor sizes only: with two factors consecutive k sit 1 + R1 = 12 points apart in the work buffer - 4-way bank conflicts
    // in the post-processing reads; measured 23.1 -> 25.0 us at 154^3 against 4.83 -> 3.55 ms at 616x1232x1232 for three factors)
    constexpr bool PFA = Pfa<R1, R2, R3>::value && R3 > 1;
    using P = Pfa<R1, R2, R3>;
    // PFA: the transformed line stays in the row-major multi-index order of the work buffer; the real-FFT post-processing looks
    // the positions of k and M - k up (table in the unused twiddle space; M + 1 <= its size for every size of the family)
    static_assert(!PFA || (TwGeom<R1, R2, R3>::TOTAL * sizeof(cx<T>) >= (M + 1) * sizeof(int)), "position table does not fit");
    int* kpos = reinterpret_cast<int*>(tw2);
    if (PFA) {
        for (int i = threadIdx.x; i <= M; i += NT) kpos[i] = P::pos_of_freq(i == M ? 0 : i);
    } else {
        build_twiddles<T, R1, R2, R3, 2>(tw2, tw3, gtw, NT);
    }
    for (int i = threadIdx.x; i < K3h; i += NT) zt[i] = gtw[i];
    if (BULK) __syncthreads();
    pdl_wait();
    if (tile < ntiles) issue(tile);
    cx<T>* a = A + l * LS;
    cx<T>* c = I + l * LS;
    for (; tile < ntiles; tile += gridDim.x) {
        if (BULK) { mbar_wait(&bar, parity); parity ^= 1; }
        else cp_async_wait_all();
        __syncthreads();
        const int L0 = tile * TL;
        const int nl = min(TL, nlines - L0);
        const bool live = l < nl;
        auto ldA = [&](int pos) { return a[pos]; };
        auto stA = [&](int pos, cx<T> v) { a[pos] = v; };
        if (PFA) {
            if (live) PStage<T, R1, 1, M / R1, JT>::run(j, P::good1z, P::good1_off, [&](int pos) { return c[pos]; }, P::rm1z, P::rm1_off, stA);
            __syncthreads();
            if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
            if (R3 > 1) {
                if (live) PStage<T, R2, 1, M / R2, JT>::run(j, P::rm2, P::rm2_off, ldA, P::rm2, P::rm2_off, stA);
                __syncthreads();
            }
            if (live) PStage<T, P::RL, 1, M / P::RL, JT>::run(j, P::rml, P::rml_off, ldA, P::rml, P::rml_off, stA);
        } else {
            fft_head<T, R1, R2, R3, 1, JT>(j, live, [&](int pos) { return c[pos]; }, stA);
            __syncthreads();
            if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
            fft_tail<T, R1, R2, R3, 1, JT, true>(j, live, tw2, tw3, ldA, stA, stA);
        }
        __syncthreads();
        cx<T>* dst = spec + (size_t)L0 * K3h;
        for (int e = threadIdx.x; e < nl * K3h; e += NT) {
            const int ll = e / K3h, k = e - ll * K3h;
            const cx<T> zk = A[ll * LS + (PFA ? kpos[k] : (k == M ? 0 : k))];
            cx<T> zc = A[ll * LS + (PFA ? kpos[M - k] : ((k == 0 || k == M) ? 0 : M - k))];
            zc.y = -zc.y;
            const cx<T> s = {(T)0.5 * (zk.x + zc.x), (T)0.5 * (zk.y + zc.y)}, d = {(T)0.5 * (zk.x - zc.x), (T)0.5 * (zk.y - zc.y)};
            const cx<T> tw = zt[k];                                   // (cos phi, -sin phi), phi = 2 pi k / K3
            const cx<T> f = {tw.y, -tw.x};                            // X = s + (-sin phi - i cos phi) d
            dst[e] = s + cmul(f, d);
        }
    }
}

// inverse: K3/2+1 complex -> K3 reals, unnormalised (tools/fft_model.py c2r)
template <typename T, int R1, int R2, int R3, int TL, int JT>
__global__ void __launch_bounds__(TL* JT, MinBlocks<TL * JT, R3>::value)
fast_z_inv_kernel(int nlines, int ntiles, const cx<T>* __restrict__ spec, T* __restrict__ mesh, const cx<T>* __restrict__ gtw, int zlc) {
    constexpr int M = R1 * R2 * R3, K3h = M + 1, LS = ZGeom<M>::LS, TILE = TL * LS, NT = TL * JT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx<T>* I = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* A = I + TILE;
    cx<T>* tw2 = A + TILE;
    cx<T>* tw3 = tw2 + TwGeom<R1, R2, R3>::N2;
    cx<T>* zt = tw3 + TwGeom<R1, R2, R3>::N3;
    // 8-line tiles: line index fastest across lanes (conflict-free shared memory, see the forward pass) and the last stage leaves the
    // real line in shared memory (natural order) for a coalesced copy-out; narrower tiles: butterfly index fastest, last stage stores
    // straight to global memory (consecutive butterflies = consecutive positions)
    constexpr bool LINE_FASTEST = TL * sizeof(cx<T>) >= 128;
    const int j = LINE_FASTEST ? threadIdx.x / TL : threadIdx.x % JT, l = LINE_FASTEST ? threadIdx.x % TL : threadIdx.x / JT;
    constexpr bool BULK = UseBulk<T>::value;
    __shared__ uint64_t bar;
    unsigned parity = 0;
    auto issue = [&](int tile) {
        const int L0 = tile * TL;
        const int nl = min(TL, nlines - L0);
        const cx<T>* src = spec + (size_t)L0 * K3h;
        if (BULK) {                                   // one bulk copy per line: K3/2 + 1 complex points
            fence_async_smem();
            if (threadIdx.x == 0) mbar_expect_tx(&bar, (unsigned)(nl * K3h * sizeof(cx<T>)));
            if (threadIdx.x < nl) bulk_g2s(I + threadIdx.x * LS, src + (size_t)threadIdx.x * K3h, (unsigned)(K3h * sizeof(cx<T>)), &bar);
        } else {
            for (int e = threadIdx.x; e < nl * K3h; e += NT) {
                const int ll = e / K3h, k = e - ll * K3h;
                cp_async<sizeof(cx<T>)>(I + ll * LS + k, src + e);
            }
            cp_async_commit();
        }
    };
    int tile = blockIdx.x;
    pdl_launch_dependents();
    if (BULK && threadIdx.x == 0) mbar_init(&bar, 1);
    build_twiddles<T, R1, R2, R3, 2>(tw2, tw3, gtw, NT);
    for (int i = threadIdx.x; i < K3h; i += NT) zt[i] = gtw[i];
    if (BULK) __syncthreads();
    pdl_wait();
    if (tile < ntiles) issue(tile);
    cx<T>* a = A + l * LS;
    cx<T>* c = I + l * LS;
    for (; tile < ntiles; tile += gridDim.x) {
        if (BULK) { mbar_wait(&bar, parity); parity ^= 1; }
        else cp_async_wait_all();
        __syncthreads();
        const int L0 = tile * TL;
        const bool live = l < nlines - L0;
        cx<T>* line = reinterpret_cast<cx<T>*>(mesh) + (size_t)(L0 + l) * zlc;
        auto untangle = [&](cx<T> xk, cx<T> xc, cx<T> tw) {           // half-complex -> packed complex: Z = s + (-sin phi + i cos phi) d
            xc.y = -xc.y;
            const cx<T> s = xk + xc, d = xk - xc;
            const cx<T> f = {tw.y, tw.x};                             // conj of the table entry gives (cos phi, +sin phi)
            return s + cmul(f, d);
        };
        // (prime-factor stages here would have to scatter 16-byte points to global memory through the CRT map: measured 0.440 -> 0.549 ms
        // at 308x616x616, 3.47 -> 4.43 ms at 616x1232x1232 - the Stockham stages keep consecutive butterflies on consecutive positions)
        fft_head<T, R1, R2, R3, -1, JT>(j, live, [&](int k) { return untangle(c[k], c[M - k], zt[k]); },
                                        [&](int pos, cx<T> v) { a[pos] = v; });
        __syncthreads();
        if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
        if (LINE_FASTEST) {
            auto ldA = [&](int pos) { return a[pos]; };
            auto stA = [&](int pos, cx<T> v) { a[pos] = v; };
            fft_tail<T, R1, R2, R3, -1, JT, true>(j, live, tw2, tw3, ldA, stA, stA);
            __syncthreads();
            const int nl = min(TL, nlines - L0);
            cx<T>* dst = reinterpret_cast<cx<T>*>(mesh) + (size_t)L0 * zlc;
            for (int e = threadIdx.x; e < nl * M; e += NT) {
                const int ll = e / M, k = e - ll * M;
                dst[(size_t)ll * zlc + k] = A[ll * LS + k];
            }
        } else {
            fft_tail<T, R1, R2, R3, -1, JT, false>(j, live, tw2, tw3, [&](int pos) { return a[pos]; }, [&](int pos, cx<T> v) { a[pos] = v; },
                                                   [&](int pos, cx<T> v) { line[pos] = v; });
        }
    }
}

// ------------------------------------------------------------------------------------------ configuration table
// (R1, R2, R3 | strided: TL, JT | z (when the size is K3/2): TL, JT); line counts are for double, float doubles TL.
// Two tile widths for the large sizes: the wide one keeps 128-byte global segments, the narrow one lets
// more blocks share an SM (shared memory: 2 * N * TL * 16 B). `pick` selects (ADMP_FFT_WIDE=1 -> wide).
#define ADMP_FAST_LIST(X)            \
    X(0, 11, 7, 1, 8, 7, 16, 7)      \
    X(1, 11, 14, 1, 8, 14, 8, 14)    \
    X(2, 11, 7, 4, 4, 28, 4, 28)     \
    X(3, 11, 7, 4, 8, 28, 8, 28)     \
    X(4, 11, 7, 8, 2, 56, 2, 56)     \
    X(5, 11, 7, 8, 4, 56, 4, 56)     \
    X(6, 11, 14, 8, 2, 112, 2, 112)  \
    X(7, 11, 7, 4, 8, 16, 8, 28)     \
    X(8, 11, 7, 8, 4, 32, 4, 56)     \
    X(9, 11, 7, 8, 8, 56, 4, 56)     \
    X(10, 11, 7, 16, 4, 112, 2, 112) \
    X(11, 11, 7, 8, 8, 56, 8, 56)

struct FastOps {
    int N, TL, threads, zTL, zthreads;
    size_t smem, smem_x, zsmem;
    int occ[6];      // strided fwd, strided inv, x conv (quick), z fwd, z inv, x conv (general)
    void (*prepare)(FastOps&);
    void (*strided)(cudaStream_t, int sign, const StrideGeom&, int ntiles, int grid, void* spec, const void* tw);
    // tensor-map variants (float64): tile loads are TMA tensor copies issued by one thread
    void (*strided_tma)(cudaStream_t, int sign, const StrideGeom&, int ntiles, int grid, void* spec, const void* tw, const CUtensorMap&);
    void (*xconv_tma)(cudaStream_t, const StrideGeom&, int tile0, int tile1, int grid, const BoxInfo*, double kappa, int kind,
                      const ConvTables&, void* spec, const void* tw, double* scalars, int want_vir, const CUtensorMap&);
    int occ_tma[4];   // strided fwd, strided inv, x conv quick, x conv general
    int box_rows;
    void (*xconv)(cudaStream_t, const StrideGeom&, int tile0, int tile1, int grid, const BoxInfo*, double kappa, int kind,
                  const ConvTables&, void* spec, const void* tw, double* scalars, int want_vir);
    // x-slab decomposed spectrum: tiles [tile0, tile1) of the X pass on peer-mapped buffers
    void (*xconv_peer)(cudaStream_t, const StrideGeom&, int tile0, int tile1, int grid, const BoxInfo*, double kappa, int kind,
                       const ConvTables&, void* spec, const void* tw, double* scalars, int want_vir, const PeerTab&, int local_reads);
    size_t smem_xp;
    int occ_peer[2];  // x conv on peers: quick, general
    size_t smem_xv;   // quick + virial variant (two more N-entry tables)
    int occ_qv[2];    // its resident blocks: cp.async tiles, TMA tiles (0: does not fit -> the general kernel runs the virial pass)
    void (*zfwd)(cudaStream_t, int nlines, int ntiles, int grid, const void* mesh, void* spec, const void* tw, int zlc);
    void (*zinv)(cudaStream_t, int nlines, int ntiles, int grid, const void* spec, void* mesh, const void* tw, int zlc);
};

template <typename K>
static int prep_kernel(K kern, int threads, size_t smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return occ;
}

template <typename T, int R1, int R2, int R3, int TLd, int JT, int ZTLd, int ZJT>
struct FastImpl {
    static constexpr int N = R1 * R2 * R3;
    static constexpr int TL = TLd * (sizeof(T) == 4 ? 2 : 1), ZTL = ZTLd * (sizeof(T) == 4 ? 2 : 1);
    static size_t smem_x_bytes(bool peer = false, bool vir = false) {
        constexpr int Q1 = R3 > 1 ? R3 : R2, Q2 = R3 > 1 ? R2 : R1, Q3 = R3 > 1 ? R1 : 1;
        return (size_t)(2 * N * TL + TwGeom<R1, R2, R3>::TOTAL + TwGeom<Q1, Q2, Q3>::TOTAL) * sizeof(cx<T>) + 2 * N * sizeof(double) +
               (peer ? N * sizeof(void*) : 0) + (vir ? 2 * N * sizeof(double) : 0);
    }
    static bool& quick_vir_ok(bool tma) {          // set by prepare(): the quick + virial instantiation fits on this device
        static bool ok[2] = {false, false};
        return ok[tma ? 1 : 0];
    }
    static void prepare(FastOps& o) {
        o.occ[0] = prep_kernel(fast_strided_kernel<T, R1, R2, R3, 1, TL, JT, false>, o.threads, o.smem);
        o.occ[1] = prep_kernel(fast_strided_kernel<T, R1, R2, R3, -1, TL, JT, false>, o.threads, o.smem);
        if (sizeof(T) == 8 && tma_box_rows(N, TL) > 0) {
            o.occ_tma[0] = prep_kernel(fast_strided_kernel<T, R1, R2, R3, 1, TL, JT, true>, o.threads, o.smem);
            o.occ_tma[1] = prep_kernel(fast_strided_kernel<T, R1, R2, R3, -1, TL, JT, true>, o.threads, o.smem);
            o.occ_tma[2] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, true>, o.threads, o.smem_x);
            o.occ_tma[3] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, false, true>, o.threads, o.smem_x);
        }
        o.occ_qv[0] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, false, true>, o.threads, o.smem_xv);
        o.occ_qv[1] = 0;
        if (sizeof(T) == 8 && tma_box_rows(N, TL) > 0)
            o.occ_qv[1] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, true, true>, o.threads, o.smem_xv);
        {
            const char* e = getenv("ADMP_FFT_QUICKVIR");
            const bool allow = !(e && atoi(e) == 0);
            quick_vir_ok(false) = allow && o.occ_qv[0] > 0;
            quick_vir_ok(true) = allow && o.occ_qv[1] > 0;
            if (!allow) o.occ_qv[0] = o.occ_qv[1] = 0;
        }
        o.occ[2] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false>, o.threads, o.smem_x);
        o.occ[5] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, false>, o.threads, o.smem_x);
        o.occ_peer[0] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, true>, o.threads, o.smem_xp);
        o.occ_peer[1] = prep_kernel(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, true>, o.threads, o.smem_xp);
        o.occ[3] = prep_kernel(fast_z_fwd_kernel<T, R1, R2, R3, ZTL, ZJT>, o.zthreads, o.zsmem);
        o.occ[4] = prep_kernel(fast_z_inv_kernel<T, R1, R2, R3, ZTL, ZJT>, o.zthreads, o.zsmem);
    }
    static void strided(cudaStream_t st, int sign, const StrideGeom& g, int ntiles, int grid, void* spec, const void* tw) {
        const size_t smem = (size_t)(2 * N * TL + TwGeom<R1, R2, R3>::TOTAL) * sizeof(cx<T>);
        const CUtensorMap none = {};
        if (sign > 0) launch_pdl(fast_strided_kernel<T, R1, R2, R3, 1, TL, JT, false>, grid, TL * JT, smem, st, g, ntiles, (cx<T>*)spec, (const cx<T>*)tw, none);
        else launch_pdl(fast_strided_kernel<T, R1, R2, R3, -1, TL, JT, false>, grid, TL * JT, smem, st, g, ntiles, (cx<T>*)spec, (const cx<T>*)tw, none);
    }
    static void strided_tma(cudaStream_t st, int sign, const StrideGeom& g, int ntiles, int grid, void* spec, const void* tw, const CUtensorMap& tm) {
        const size_t smem = (size_t)(2 * N * TL + TwGeom<R1, R2, R3>::TOTAL) * sizeof(cx<T>);
        if (sign > 0) launch_pdl(fast_strided_kernel<T, R1, R2, R3, 1, TL, JT, true>, grid, TL * JT, smem, st, g, ntiles, (cx<T>*)spec, (const cx<T>*)tw, tm);
        else launch_pdl(fast_strided_kernel<T, R1, R2, R3, -1, TL, JT, true>, grid, TL * JT, smem, st, g, ntiles, (cx<T>*)spec, (const cx<T>*)tw, tm);
    }
    static void xconv_tma(cudaStream_t st, const StrideGeom& g, int tile0, int ntiles, int grid, const BoxInfo* B, double kappa, int kind,
                          const ConvTables& tb, void* spec, const void* tw, double* scalars, int want_vir, const CUtensorMap& tm) {
        const size_t smem = smem_x_bytes();
        if (kind == ADMP_CK_COULOMB && !want_vir)
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, true>, grid, TL * JT, smem, st, g, tile0, ntiles, B, (T)kappa, kind, tb,
                       (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, tm);
        else if (kind == ADMP_CK_COULOMB && quick_vir_ok(true))
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, true, true>, grid, TL * JT, smem_x_bytes(false, true), st, g, tile0, ntiles, B,
                       (T)kappa, kind, tb, (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, tm);
        else
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, false, true>, grid, TL * JT, smem, st, g, tile0, ntiles, B, (T)kappa, kind, tb,
                       (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, tm);
    }
    static void xconv(cudaStream_t st, const StrideGeom& g, int tile0, int ntiles, int grid, const BoxInfo* B, double kappa, int kind,
                      const ConvTables& tb, void* spec, const void* tw, double* scalars, int want_vir) {
        const size_t smem = smem_x_bytes();
        if (kind == ADMP_CK_COULOMB && !want_vir)
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false>, grid, TL * JT, smem, st, g, tile0, ntiles, B, (T)kappa, kind, tb,
                       (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, CUtensorMap{});
        else if (kind == ADMP_CK_COULOMB && quick_vir_ok(false))
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, false, false, true>, grid, TL * JT, smem_x_bytes(false, true), st, g, tile0, ntiles, B,
                       (T)kappa, kind, tb, (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, CUtensorMap{});
        else
            launch_pdl(fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, false>, grid, TL * JT, smem, st, g, tile0, ntiles, B, (T)kappa, kind, tb,
                       (cx<T>*)spec, (const cx<T>*)tw, scalars, want_vir, PeerTab{}, 0, CUtensorMap{});
    }
    static void xconv_peer(cudaStream_t st, const StrideGeom& g, int tile0, int tile1, int grid, const BoxInfo* B, double kappa, int kind,
                           const ConvTables& tb, void* spec, const void* tw, double* scalars, int want_vir, const PeerTab& peers,
                           int local_reads) {
        const size_t smem = smem_x_bytes(true);
        if (kind == ADMP_CK_COULOMB && !want_vir)
            fast_x_conv_kernel<T, R1, R2, R3, TL, JT, true, true><<<grid, TL * JT, smem, st>>>(g, tile0, tile1, B, (T)kappa, kind, tb, (cx<T>*)spec,
                                                                                             (const cx<T>*)tw, scalars, want_vir, peers, local_reads,
                                                                                             CUtensorMap{});
        else
            fast_x_conv_kernel<T, R1, R2, R3, TL, JT, false, true><<<grid, TL * JT, smem, st>>>(g, tile0, tile1, B, (T)kappa, kind, tb, (cx<T>*)spec,
                                                                                              (const cx<T>*)tw, scalars, want_vir, peers, local_reads,
                                                                                              CUtensorMap{});
    }
    static void zfwd(cudaStream_t st, int nlines, int ntiles, int grid, const void* mesh, void* spec, const void* tw, int zlc) {
        const size_t smem = (size_t)(2 * ZTL * ZGeom<N>::LS + TwGeom<R1, R2, R3>::TOTAL + N + 1) * sizeof(cx<T>);
        launch_pdl(fast_z_fwd_kernel<T, R1, R2, R3, ZTL, ZJT>, grid, ZTL * ZJT, smem, st, nlines, ntiles, (const T*)mesh, (cx<T>*)spec, (const cx<T>*)tw, zlc);
    }
    static void zinv(cudaStream_t st, int nlines, int ntiles, int grid, const void* spec, void* mesh, const void* tw, int zlc) {
        const size_t smem = (size_t)(2 * ZTL * ZGeom<N>::LS + TwGeom<R1, R2, R3>::TOTAL + N + 1) * sizeof(cx<T>);
        launch_pdl(fast_z_inv_kernel<T, R1, R2, R3, ZTL, ZJT>, grid, ZTL * ZJT, smem, st, nlines, ntiles, (const cx<T>*)spec, (T*)mesh, (const cx<T>*)tw, zlc);
    }
    static FastOps ops() {
        FastOps o = {};
        o.N = N; o.TL = TL; o.threads = TL * JT; o.zTL = ZTL; o.zthreads = ZTL * ZJT;
        o.smem = (size_t)(2 * N * TL + TwGeom<R1, R2, R3>::TOTAL) * sizeof(cx<T>);
        o.smem_x = smem_x_bytes();
        o.smem_xp = smem_x_bytes(true);
        o.smem_xv = smem_x_bytes(false, true);
        o.zsmem = (size_t)(2 * ZTL * ZGeom<N>::LS + TwGeom<R1, R2, R3>::TOTAL + N + 1) * sizeof(cx<T>);
        o.box_rows = tma_box_rows(N, TL);
        o.strided_tma = &strided_tma; o.xconv_tma = &xconv_tma;
        o.prepare = &prepare; o.strided = &strided; o.xconv = &xconv; o.xconv_peer = &xconv_peer; o.zfwd = &zfwd; o.zinv = &zinv;
        return o;
    }
};

// the table entries for (size N, element size); wide = prefer the wider tile when two are listed
// entries 7, 9 and 11 (11: 8-line Z tiles of 616 points for the forward Z pass) are only taken when asked for by id (ADMP_FFT_XCFG / YCFG / ZFCFG); 8 (616 points in 128-thread blocks, two
// per SM) is the wide default of its size: Y pass 0.380 -> 0.362 ms at 308x616x616, X pass 7.04 -> 6.74 ms at 616x1232x1232
template <typename T>
static bool fast_lookup(int N, bool wide, FastOps& out, int force_id = -1) {
    bool found = false;
#define X(id, a, b, c, tl, jt, ztl, zjt)                                                   \
    if (N == (a) * (b) * (c) && force_id == (id)) { out = FastImpl<T, a, b, c, tl, jt, ztl, zjt>::ops(); return true; }
    ADMP_FAST_LIST(X)
#undef X
#define X(id, a, b, c, tl, jt, ztl, zjt)                                                   \
    if (((id) < 7 || (id) == 8 || (id) == 10) && N == (a) * (b) * (c) && (!found || wide)) { out = FastImpl<T, a, b, c, tl, jt, ztl, zjt>::ops(); found = true; }
    ADMP_FAST_LIST(X)
#undef X
    return found;
}
