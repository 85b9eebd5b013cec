// Per-site terms and the device-side control of the induced-dipole loop.
//   * box_setup_kernel   : BoxInfo (inverse, Nstar, volume) from the caller's box
//   * self_kernel        : pme_self + pol_penalty (admp/pme.py:738-774) with adjoints
//   * scf_* kernels      : optimize_Uind's convergence test / Jacobi update
//                          (admp/pme.py:126-143) without a host round trip
#include "kernels.h"

namespace admp {

template <typename T>
__global__ void box_setup_kernel(const T* __restrict__ box, BoxInfo* __restrict__ B, int K1, int K2, int K3) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double b[9];
    for (int k = 0; k < 9; ++k) { b[k] = (double)box[k]; B->box[k] = b[k]; }
    const double det = b[0] * (b[4] * b[8] - b[5] * b[7]) - b[1] * (b[3] * b[8] - b[5] * b[6]) + b[2] * (b[3] * b[7] - b[4] * b[6]);
    const double id = 1.0 / det;
    double inv[9];
    inv[0] = (b[4] * b[8] - b[5] * b[7]) * id; inv[1] = (b[2] * b[7] - b[1] * b[8]) * id; inv[2] = (b[1] * b[5] - b[2] * b[4]) * id;
    inv[3] = (b[5] * b[6] - b[3] * b[8]) * id; inv[4] = (b[0] * b[8] - b[2] * b[6]) * id; inv[5] = (b[2] * b[3] - b[0] * b[5]) * id;
    inv[6] = (b[3] * b[7] - b[4] * b[6]) * id; inv[7] = (b[1] * b[6] - b[0] * b[7]) * id; inv[8] = (b[0] * b[4] - b[1] * b[3]) * id;
    const int K[3] = {K1, K2, K3};
    for (int k = 0; k < 9; ++k) B->inv[k] = inv[k];
    for (int d = 0; d < 3; ++d) {
        B->K[d] = K[d];
        for (int c = 0; c < 3; ++c) B->nstar[3 * d + c] = (double)K[d] * inv[3 * c + d];      // recip.py:55
    }
    B->vol = det;
}

// E_self = -DIEL sum_l kappa/sqrt(pi) (2 kappa^2)^l/(2l+1)!! |Q_l|^2 with |Q_2|^2 = (2/3) T:T,
// E_pen = DIEL sum U^2 / (2 max(pol,1e-8)). mu_tot = mu + U enters the dipole term (pme.py:233-252).
template <typename T>
__global__ void __launch_bounds__(128)
self_kernel(int n, T kappa, const T* __restrict__ M, const T* __restrict__ U, const T* __restrict__ pol, uint32_t flags,
            T* __restrict__ G, T* __restrict__ F, T* __restrict__ dpol, double* __restrict__ scalars) {
    __shared__ double red[2 * 4];
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[2] = {0.0, 0.0};
    if (a < n) {
        const double k = (double)kappa, D = ADMP_DIEL;
        const double f0 = k / ADMP_SQRT_PI, f1 = f0 * (2 * k * k) / 3, f2 = f0 * (2 * k * k) * (2 * k * k) / 15;
        const T* m = M + (size_t)a * 10;
        double mu[3] = {(double)m[1], (double)m[2], (double)m[3]};
        double u[3] = {0, 0, 0};
        if (U != nullptr) { for (int c = 0; c < 3; ++c) { u[c] = (double)U[3 * a + c]; mu[c] += u[c]; } }
        const double q = (double)m[0];
        const double tt = (double)m[4] * m[4] + (double)m[7] * m[7] + (double)m[9] * m[9]
                        + 2 * ((double)m[5] * m[5] + (double)m[6] * m[6] + (double)m[8] * m[8]);
        acc[0] = -D * (f0 * q * q + f1 * (mu[0] * mu[0] + mu[1] * mu[1] + mu[2] * mu[2]) + f2 * (2.0 / 3) * tt);
        double pt = 1.0, pen_g[3] = {0, 0, 0};
        if (U != nullptr) {
            const double p = (double)pol[a];
            pt = p < 1e-8 ? 1e-8 : p;                                   // trim_val_0, pme.py:771
            const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
            acc[1] = D * 0.5 / pt * uu;
            for (int c = 0; c < 3; ++c) pen_g[c] = D * u[c] / pt;
            if ((flags & ADMP_WANT_PGRAD) && dpol != nullptr && p >= 1e-8) atomicAdd(dpol + a, (T)(-D * 0.5 * uu / (pt * pt)));
        }
        if (flags & ADMP_WANT_GRAD) {
            if (G != nullptr) {
                T* g = G + (size_t)a * 10;
                atomicAdd(g, (T)(-2 * D * f0 * q));
                for (int c = 0; c < 3; ++c) atomicAdd(g + 1 + c, (T)(-2 * D * f1 * mu[c]));
                const double c1 = -D * f2 * (4.0 / 3), c2 = -D * f2 * (8.0 / 3);
                atomicAdd(g + 4, (T)(c1 * m[4])); atomicAdd(g + 5, (T)(c2 * m[5])); atomicAdd(g + 6, (T)(c2 * m[6]));
                atomicAdd(g + 7, (T)(c1 * m[7])); atomicAdd(g + 8, (T)(c2 * m[8])); atomicAdd(g + 9, (T)(c1 * m[9]));
            }
            if (F != nullptr && U != nullptr)
                for (int c = 0; c < 3; ++c) atomicAdd(F + 3 * a + c, (T)(-2 * D * f1 * mu[c] + pen_g[c]));
        }
    }
    block_accumulate<2>(acc, red, scalars + ADMP_S_E_SELF);     // E_SELF, E_PEN are adjacent slots
}

// dispersion self term, admp/disp_pme.py:254-279
template <typename T>
__global__ void __launch_bounds__(128)
disp_self_kernel(int n, T kappa, int pmax, const T* __restrict__ c_list, uint32_t flags, T* __restrict__ dc,
                 double* __restrict__ scalars) {
    __shared__ double red[4];
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0.0};
    if (a < n) {
        const double k2 = (double)kappa * kappa, k6 = k2 * k2 * k2;
        const double f[3] = {-k6 / 12, -k6 * k2 / 48, -k6 * k2 * k2 / 240};
        const int np = (pmax - 4) / 2;
        for (int p = 0; p < 3 && p < np; ++p) {
            const double c = (double)c_list[(size_t)a * 3 + p];
            acc[0] += f[p] * c * c;
            if ((flags & ADMP_WANT_PGRAD) && dc != nullptr) atomicAdd(dc + (size_t)a * 3 + p, (T)(2 * f[p] * c));
        }
    }
    block_accumulate<1>(acc, red, scalars + ADMP_S_E_SELF);
}

// SCF: F += self/penalty part of dE/dU; reduce max |F| over sites with pol > 0.001 (pme.py:125,136)
template <typename T>
__global__ void __launch_bounds__(128)
scf_field_kernel(int n, T kappa, const T* __restrict__ M, const T* __restrict__ U, const T* __restrict__ pol,
                 T* __restrict__ F, double* __restrict__ scalars) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (a < n) {
        const double k = (double)kappa, D = ADMP_DIEL;
        const double f1 = k / ADMP_SQRT_PI * (2 * k * k) / 3;
        const double p = (double)pol[a], pt = p < 1e-8 ? 1e-8 : p;
        for (int c = 0; c < 3; ++c) {
            const double u = (double)U[3 * a + c];
            const double f = (double)F[3 * a + c] - 2 * D * f1 * ((double)M[(size_t)a * 10 + 1 + c] + u) + D * u / pt;
            F[3 * a + c] = (T)f;
            // a NaN / Inf field must read as "not converged" (the reference's `NaN < thresh` is False): fmax would drop it
            if (p > 0.001) mx = (f == f && fabs(f) <= 1.7e308) ? fmax(mx, fabs(f)) : 1.7976931348623157e308;
        }
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomic_max_nonneg(scalars + ADMP_S_MAXFIELD, mx);
}

// Decision step of optimize_Uind (pme.py:133-143): test BEFORE the update; converging exactly on
// the last iteration reports flag False. state: [0]=iter, [1]=do_update, [2]=final_pass,
// [3]=n_cycle, [4]=converged, [5]=loop condition (mirrors the graph conditional handle), [7]=loop ended.
__global__ void scf_decide_kernel(int32_t* __restrict__ state, double* __restrict__ scalars, int maxiter, double thresh,
                                  cudaGraphConditionalHandle handle, int use_handle, int refresh_in_loop) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (state[7]) {                       // the loop has already ended: a speculatively enqueued extra cycle (host-driven
        state[1] = 0;                     // multi-GPU loop, admp_b200/parallel.py) must not touch U or the result
        state[5] = 0;
        if (use_handle) cudaGraphSetConditional(handle, 0);
        return;
    }
    int cond = 0;
    if (state[2]) {                       // this pass only refreshed the field/mesh after the last update
        state[1] = 0;
    } else {
        const int it = state[0];
        const double mx = __longlong_as_double((long long)reinterpret_cast<unsigned long long*>(scalars)[ADMP_S_MAXFIELD]);
        if (mx < thresh) {
            state[1] = 0; state[3] = it; state[4] = (it >= maxiter - 1) ? 0 : 1;
        } else {
            state[1] = 1;
            if (it >= maxiter - 1) {
                // last allowed cycle: U is still updated (pme.py:138), flag False. The mesh of the updated U is
                // rebuilt by one more pass of the loop body, or by the caller's final (virial) pass.
                state[3] = it; state[4] = 0;
                if (refresh_in_loop) { state[2] = 1; cond = 1; }
            } else {
                state[0] = it + 1;
                cond = 1;
            }
        }
    }
    state[5] = cond;
    if (!cond) state[7] = 1;
    if (cond) {                           // re-arm the per-cycle accumulators for the next pass
        scalars[ADMP_S_MAXFIELD] = 0.0;
        scalars[ADMP_S_E_RECIP] = 0.0;
    }
    if (use_handle) cudaGraphSetConditional(handle, cond);
}

// Jacobi update U <- U - F pol / DIEL (pme.py:138) when the decision says so; re-arms the
// per-iteration accumulators either way.
template <typename T>
__global__ void __launch_bounds__(128)
scf_update_kernel(int n, const int32_t* __restrict__ state, T* __restrict__ F, const T* __restrict__ pol,
                  T* __restrict__ U, int zero_F) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    if (state[1]) {
        const T s = pol[a] * (T)(1.0 / ADMP_DIEL);
        for (int c = 0; c < 3; ++c) U[3 * a + c] -= F[3 * a + c] * s;
    }
    if (zero_F) {                         // the next cycle accumulates its field with atomics
        for (int c = 0; c < 3; ++c) F[3 * a + c] = (T)0;
    }
}

// ---- beyond the reference: conjugate gradients on the same fixed point, preconditioned with the
// Jacobi scaling pol / DIEL.  The loop body is the Jacobi one (pair field + reciprocal field + scf_field_kernel, all
// evaluated on the dipoles in `Us`); this kernel replaces decide + update.  One block: the dot products and the maximum
// are block reductions in a fixed order, so a run is reproducible bit for bit.
//   phase 0 (state[6] = 0): F = field(U), Us = U.  max|F| (scalars[ADMP_S_MAXFIELD], sites with pol > 0.001) < thresh
//            -> converged, stop: the mesh and the reciprocal energy of this pass belong to the final U.  Else
//            r = -F, z = M r, p = z, Us = U + p.
//   phase 1 (state[6] = 1): F = field(U + p), so A p = F + r;  alpha = rz / pAp;  U += alpha p;  r -= alpha A p;
//            max|r| < thresh or the iteration budget is spent -> Us = U, back to phase 0 (the TRUE residual decides);
//            else z = M r, beta = rz' / rz, p = z + beta p, Us = U + p.
//            pAp <= 0 (A not positive definite along p): Us = U, one refresh pass, stop with flag 0.
// cg = [U | r | p] (3n doubles each) + 1 double (rz).  state as in scf_decide_kernel; [0] counts CG iterations.
__device__ __forceinline__ double cg_block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}
__device__ __forceinline__ double cg_block_max(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = fmax(s, red[w]);
    return s;
}
template <typename T>
__global__ void __launch_bounds__(1024)
scf_cg_kernel(int n, int32_t* __restrict__ state, double* __restrict__ scalars, T* __restrict__ F, const T* __restrict__ pol,
              T* __restrict__ Us, double* __restrict__ cg, int maxiter, double thresh, cudaGraphConditionalHandle handle, int use_handle) {
    __shared__ double red[32];
    const int n3 = 3 * n;
    double* U = cg;
    double* r = cg + n3;
    double* p = cg + 2 * (size_t)n3;
    double* rzp = cg + 3 * (size_t)n3;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double invD = 1.0 / ADMP_DIEL;
    auto finish = [&](int cond) {           // every thread calls it (uniform `cond`)
        for (int e = tid; e < n3; e += nt) F[e] = (T)0;          // the next pass accumulates its field with atomics
        if (tid == 0) {
            state[5] = cond;
            if (!cond) state[7] = 1;
            if (cond) {
                scalars[ADMP_S_MAXFIELD] = 0.0;
                scalars[ADMP_S_E_RECIP] = 0.0;
            }
            if (use_handle) cudaGraphSetConditional(handle, cond);
        }
    };
    if (state[7]) { finish(0); return; }     // the loop has ended (speculative extra pass): leave everything alone
    if (state[2]) { finish(0); return; }     // this pass only refreshed the mesh / field on the final U
    const int it = state[0];
    if (state[6] == 0) {
        const double mx = __longlong_as_double((long long)reinterpret_cast<unsigned long long*>(scalars)[ADMP_S_MAXFIELD]);
        if (mx < thresh || it >= maxiter) {
            if (tid == 0) { state[3] = it; state[4] = (mx < thresh) ? 1 : 0; }
            finish(0);
            return;
        }
        double acc = 0.0;
        for (int e = tid; e < n3; e += nt) {
            const double u = (double)Us[e], rr = -(double)F[e], z = (double)pol[e / 3] * invD * rr;
            U[e] = u; r[e] = rr; p[e] = z;
            Us[e] = (T)(u + z);
            acc = fma(rr, z, acc);
        }
        const double rz = cg_block_sum(acc, red);
        if (tid == 0) { *rzp = rz; state[6] = 1; }
        finish(1);
        return;
    }
    // phase 1
    double acc = 0.0;
    for (int e = tid; e < n3; e += nt) acc = fma(p[e], (double)F[e] + r[e], acc);
    const double pAp = cg_block_sum(acc, red);
    const double rz = *rzp;
    if (!(pAp > 0.0) || !(rz == rz)) {       // indefinite direction (or NaN): stop on the current U, flag 0
        for (int e = tid; e < n3; e += nt) Us[e] = (T)U[e];
        __syncthreads();
        if (tid == 0) { state[3] = it; state[4] = 0; state[2] = 1; state[6] = 0; }
        finish(1);
        return;
    }
    const double alpha = rz / pAp;
    double mx = 0.0;
    acc = 0.0;
    for (int e = tid; e < n3; e += nt) {
        const double pe = p[e], pl = (double)pol[e / 3];
        const double rr = r[e] - alpha * ((double)F[e] + r[e]);
        U[e] += alpha * pe;
        r[e] = rr;
        if (pl > 0.001) mx = (rr == rr && fabs(rr) <= 1.7e308) ? fmax(mx, fabs(rr)) : 1.7976931348623157e308;
        acc = fma(rr, pl * invD * rr, acc);
    }
    const double rz_new = cg_block_sum(acc, red);
    mx = cg_block_max(mx, red);
    if (mx < thresh || it + 1 >= maxiter) {
        for (int e = tid; e < n3; e += nt) Us[e] = (T)U[e];
        __syncthreads();
        if (tid == 0) { state[0] = it + 1; state[6] = 0; }
        finish(1);
        return;
    }
    const double beta = rz_new / rz;
    for (int e = tid; e < n3; e += nt) {
        const double pe = (double)pol[e / 3] * invD * r[e] + beta * p[e];
        p[e] = pe;
        Us[e] = (T)(U[e] + pe);
    }
    __syncthreads();
    if (tid == 0) { state[0] = it + 1; *rzp = rz_new; }
    finish(1);
}

// dE/dbox (3x3, row-major) += reciprocal-space part assembled from the accumulators:
//   -inv^T (W^T Nstar)  [spread/gather, W = dE/dNstar]  - 2 T inv^T - E_recip inv^T  [k^2 and V in C_k]
__global__ void virial_finalize_kernel(const BoxInfo* __restrict__ B, double* __restrict__ s, int kvec_ref) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* W = s + ADMP_S_DNSTAR;
    const double* tk = s + ADMP_S_TK;
    const double Tn[9] = {tk[0], tk[1], tk[2], tk[1], tk[3], tk[4], tk[2], tk[4], tk[5]};
    double T[9];
    // kvec_ref: the reference's k table is built with meshgrid(kz, kx, ky) (admp/recip.py:339-341), which pairs
    // k-vector component 0 with mesh axis 1 and vice versa: T_ref[a][b] = T[p(a)][p(b)], p = (1, 0, 2).
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            const int pa = kvec_ref ? (a == 0 ? 1 : (a == 1 ? 0 : 2)) : a, pb = kvec_ref ? (b == 0 ? 1 : (b == 1 ? 0 : 2)) : b;
            T[3 * a + b] = Tn[3 * pa + pb];
        }
    const double E = s[ADMP_S_E_RECIP];
    double Mm[9];
    for (int c = 0; c < 3; ++c)
        for (int b = 0; b < 3; ++b) Mm[3 * c + b] = W[c] * B->nstar[b] + W[3 + c] * B->nstar[3 + b] + W[6 + c] * B->nstar[6 + b];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double v = 0.0;
            for (int c = 0; c < 3; ++c) v += -B->inv[3 * c + a] * Mm[3 * c + b] - 2.0 * T[3 * a + c] * B->inv[3 * b + c];
            v -= E * B->inv[3 * b + a];
            s[ADMP_S_DBOX + 3 * a + b] += v;
        }
}
void launch_virial_finalize(cudaStream_t st, const BoxInfo* B, double* scalars, int kvec_ref) {
    virial_finalize_kernel<<<1, 32, 0, st>>>(B, scalars, kvec_ref);
}

template <typename T>
void launch_box_setup(cudaStream_t st, const void* box, BoxInfo* B, int K1, int K2, int K3) {
    box_setup_kernel<T><<<1, 32, 0, st>>>((const T*)box, B, K1, K2, K3);
}
template <typename T>
void launch_self(cudaStream_t st, int n, double kappa, const void* M, const void* U, const void* pol, uint32_t flags, void* G,
                 void* F, void* dpol, double* scalars) {
    if (n <= 0) return;
    self_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, (T)kappa, (const T*)M, (const T*)U, (const T*)pol, flags, (T*)G, (T*)F, (T*)dpol, scalars);
}
template <typename T>
void launch_disp_self(cudaStream_t st, int n, double kappa, int pmax, const void* c_list, uint32_t flags, void* dc, double* scalars) {
    if (n <= 0) return;
    disp_self_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, (T)kappa, pmax, (const T*)c_list, flags, (T*)dc, scalars);
}
template <typename T>
void launch_scf_field(cudaStream_t st, int n, double kappa, const void* M, const void* U, const void* pol, void* F, double* scalars) {
    scf_field_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, (T)kappa, (const T*)M, (const T*)U, (const T*)pol, (T*)F, scalars);
}
void launch_scf_decide(cudaStream_t st, int32_t* state, double* scalars, int maxiter, double thresh, int refresh_in_loop,
                       cudaGraphConditionalHandle handle, int use_handle) {
    scf_decide_kernel<<<1, 32, 0, st>>>(state, scalars, maxiter, thresh, handle, use_handle, refresh_in_loop);
}
template <typename T>
void launch_scf_cg(cudaStream_t st, int n, int32_t* state, double* scalars, void* F, const void* pol, void* Us, double* cg, int maxiter,
                   double thresh, cudaGraphConditionalHandle handle, int use_handle) {
    scf_cg_kernel<T><<<1, 1024, 0, st>>>(n, state, scalars, (T*)F, (const T*)pol, (T*)Us, cg, maxiter, thresh, handle, use_handle);
}
template <typename T>
void launch_scf_update(cudaStream_t st, int n, const int32_t* state, void* F, const void* pol, void* U, int zero_F) {
    scf_update_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, state, (T*)F, (const T*)pol, (T*)U, zero_F);
}

#define ADMP_INST(T)                                                                                                         \
    template void launch_box_setup<T>(cudaStream_t, const void*, BoxInfo*, int, int, int);                                   \
    template void launch_self<T>(cudaStream_t, int, double, const void*, const void*, const void*, uint32_t, void*, void*,   \
                                 void*, double*);                                                                            \
    template void launch_disp_self<T>(cudaStream_t, int, double, int, const void*, uint32_t, void*, double*);                \
    template void launch_scf_field<T>(cudaStream_t, int, double, const void*, const void*, const void*, void*, double*);     \
    template void launch_scf_update<T>(cudaStream_t, int, const int32_t*, void*, const void*, void*, int);                   \
    template void launch_scf_cg<T>(cudaStream_t, int, int32_t*, double*, void*, const void*, void*, double*, int, double,   \
                                   cudaGraphConditionalHandle, int);
ADMP_INST(double)
ADMP_INST(float)
#undef ADMP_INST

}  // namespace admp
