// Per-site kernels: local frames + local->global rotation of multipoles, and the adjoint
// (torque -> anchor-atom forces, dE/dQ_local, image-shift box term).
//
// Replaces admp/spatial.py:44-147 (generate_construct_local_frames) and
// admp/multipole.py:92-201 (rot_local2global) and, for the backward pass, what jax.grad
// derives from them. One thread per site; HBM-bound: Na*(3*3 + 9 + 10) reals.
#include "kernels.h"

namespace admp {

enum { ZTHENX = 0, BISECTOR = 1, ZBISECT = 2, THREEFOLD = 3, ZONLY = 4, NOAXIS = 5 };

template <typename T> struct V3 { T x, y, z; };
template <typename T> __device__ __forceinline__ V3<T> mk(T x, T y, T z) { V3<T> r = {x, y, z}; return r; }
template <typename T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator*(V3<T> a, T s) { return mk(a.x * s, a.y * s, a.z * s); }
template <typename T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// u = v/|v| ; adjoint: gv = (gu - u (u.gu)) / |v|
template <typename T> __device__ __forceinline__ V3<T> unit(V3<T> v, T& nrm) { nrm = sqrt(dot(v, v)); return v * ((T)1 / nrm); }
template <typename T> __device__ __forceinline__ V3<T> unit_bwd(V3<T> g, V3<T> u, T nrm) { return (g - u * dot(u, g)) * ((T)1 / nrm); }

template <typename T>
__device__ __forceinline__ V3<T> anchor_vec(const BoxInfo& B, const T* __restrict__ pos, int a, int b, T (&sh)[3]) {
    T d[3] = {pos[3 * b] - pos[3 * a], pos[3 * b + 1] - pos[3 * a + 1], pos[3 * b + 2] - pos[3 * a + 2]};
    min_image(B, d, sh);
    return mk(d[0], d[1], d[2]);
}

// everything the backward pass needs from the frame construction
template <typename T> struct FrameState {
    V3<T> vz0, vx0, vy0, vz, vx, vxp, vy;
    T nz, nx, ny, nzb, nxb, nw, s;
    T shz[3], shx[3], shy[3];
    int type, az, ax, ay;
};

template <typename T>
__device__ __forceinline__ void build_frame(const BoxInfo& B, const T* __restrict__ pos, int a, int type,
                                            const int32_t* __restrict__ ai, FrameState<T>& f) {
    f.type = type;
    f.az = ai[3 * a]; f.ax = ai[3 * a + 1]; f.ay = ai[3 * a + 2];
    f.vz0 = unit(anchor_vec(B, pos, a, f.az, f.shz), f.nz);                       // spatial.py:98-99
    if (type == ZONLY) {                                                          // :103-105
        T xz0 = rint(fabs(f.vz0.x));
        f.vx0 = mk((T)1 - xz0, xz0, (T)0);
        f.nx = (T)1;
    } else {
        f.vx0 = unit(anchor_vec(B, pos, a, f.ax, f.shx), f.nx);                   // :107-110
    }
    f.vz = f.vz0; f.vx = f.vx0;
    if (type == ZBISECT || type == THREEFOLD) f.vy0 = unit(anchor_vec(B, pos, a, f.ay, f.shy), f.ny);
    if (type == BISECTOR) f.vz = unit(f.vz0 + f.vx0, f.nzb);                      // :112-114
    if (type == ZBISECT) f.vx = unit(f.vx0 + f.vy0, f.nxb);                       // :116-121
    if (type == THREEFOLD) f.vz = unit(f.vz0 + f.vx0 + f.vy0, f.nzb);             // :123-135
    f.s = dot(f.vx, f.vz);                                                        // :138-139
    f.vxp = unit(f.vx - f.vz * f.s, f.nw);
    f.vy = cross(f.vz, f.vxp);                                                    // :141
}

// harmonic (9) -> local Cartesian layout (10)
template <typename T>
__device__ __forceinline__ void harm_to_cart(const T* __restrict__ Q, int lmax, T (&m)[10]) {
    const T h = (T)(ADMP_SQRT3 / 2);
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = (T)0;
    m[0] = Q[0];
    if (lmax >= 1) { m[1] = Q[2]; m[2] = Q[3]; m[3] = Q[1]; }
    if (lmax >= 2) {
        m[4] = (T)-0.5 * Q[4] + h * Q[7];
        m[5] = h * Q[8];
        m[6] = h * Q[5];
        m[7] = (T)-0.5 * Q[4] - h * Q[7];
        m[8] = h * Q[6];
        m[9] = Q[4];
    }
}
template <typename T>
__device__ __forceinline__ void cart_to_harm(const T (&m)[10], T* __restrict__ Q, int nh) {
    const T c = (T)(2 / ADMP_SQRT3);
    Q[0] = m[0];
    if (nh >= 4) { Q[1] = m[3]; Q[2] = m[1]; Q[3] = m[2]; }
    if (nh >= 9) {
        Q[4] = m[9]; Q[5] = c * m[6]; Q[6] = c * m[8];
        Q[7] = (m[4] - m[7]) * (T)(1 / ADMP_SQRT3); Q[8] = c * m[5];
    }
}
// adjoint of harm_to_cart
template <typename T>
__device__ __forceinline__ void cart_grad_to_harm(const T (&g)[10], T* __restrict__ dQ, int nh) {
    const T h = (T)(ADMP_SQRT3 / 2);
    dQ[0] = g[0];
    if (nh >= 4) { dQ[2] = g[1]; dQ[3] = g[2]; dQ[1] = g[3]; }
    if (nh >= 9) {
        dQ[4] = g[9] - (T)0.5 * g[4] - (T)0.5 * g[7];
        dQ[5] = h * g[6]; dQ[6] = h * g[8]; dQ[7] = h * (g[4] - g[7]); dQ[8] = h * g[5];
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
frames_fwd_kernel(int n, int lmax, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos,
                  const int32_t* __restrict__ atype, const int32_t* __restrict__ ai,
                  const T* __restrict__ Ql, T* __restrict__ M, T* __restrict__ Qg, T* __restrict__ Fr) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int nh = (lmax + 1) * (lmax + 1);
    T ml[10];
    harm_to_cart(Ql + (size_t)a * nh, lmax, ml);
    const int type = (atype != nullptr && lmax > 0) ? atype[a] : NOAXIS;
    T R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (type != NOAXIS) {
        FrameState<T> f;
        build_frame(*Bp, pos, a, type, ai, f);
        R[0] = f.vxp.x; R[1] = f.vxp.y; R[2] = f.vxp.z;
        R[3] = f.vy.x;  R[4] = f.vy.y;  R[5] = f.vy.z;
        R[6] = f.vz.x;  R[7] = f.vz.y;  R[8] = f.vz.z;
    }
    // global = R^T local  (rot_local2global == rot_global2local with the transposed frame)
    T mg[10];
    mg[0] = ml[0];
#pragma unroll
    for (int c = 0; c < 3; ++c) mg[1 + c] = R[c] * ml[1] + R[3 + c] * ml[2] + R[6 + c] * ml[3];
    const T Tl[9] = {ml[4], ml[5], ml[6], ml[5], ml[7], ml[8], ml[6], ml[8], ml[9]};
    T TR[9];   // TR[i][b] = sum_j Tl[i][j] R[j][b]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int b = 0; b < 3; ++b) TR[3 * i + b] = Tl[3 * i] * R[b] + Tl[3 * i + 1] * R[3 + b] + Tl[3 * i + 2] * R[6 + b];
    auto tg = [&](int p, int b) { return R[p] * TR[b] + R[3 + p] * TR[3 + b] + R[6 + p] * TR[6 + b]; };
    mg[4] = tg(0, 0); mg[5] = tg(0, 1); mg[6] = tg(0, 2); mg[7] = tg(1, 1); mg[8] = tg(1, 2); mg[9] = tg(2, 2);
    if (M != nullptr) {
#pragma unroll
        for (int k = 0; k < 10; ++k) M[(size_t)a * 10 + k] = mg[k];
    }
    if (Qg != nullptr) cart_to_harm(mg, Qg + (size_t)a * nh, nh);
    if (Fr != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Fr[(size_t)a * 9 + k] = R[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
frames_bwd_kernel(int first, int n, int lmax, const BoxInfo* __restrict__ Bp, const T* __restrict__ pos,
                  const int32_t* __restrict__ atype, const int32_t* __restrict__ ai,
                  const T* __restrict__ Ql, const T* __restrict__ G, T* __restrict__ dQl,
                  T* __restrict__ dpos, double* __restrict__ scalars, int want_box) {
    __shared__ double red[9 * 4];
    const int a = first + blockIdx.x * blockDim.x + threadIdx.x;     // atoms [first, n)
    double dbox[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (a < n) {
        const int nh = (lmax + 1) * (lmax + 1);
        T g[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) g[k] = G[(size_t)a * 10 + k];
        const int type = (atype != nullptr && lmax > 0) ? atype[a] : NOAXIS;
        if (type == NOAXIS) {
            cart_grad_to_harm(g, dQl + (size_t)a * nh, nh);
        } else {
            T ml[10];
            harm_to_cart(Ql + (size_t)a * nh, lmax, ml);
            FrameState<T> f;
            build_frame(*Bp, pos, a, type, ai, f);
            const T R[9] = {f.vxp.x, f.vxp.y, f.vxp.z, f.vy.x, f.vy.y, f.vy.z, f.vz.x, f.vz.y, f.vz.z};
            // symmetric matrix form of the quadrupole gradient (half of the off-diagonals)
            const T Gs[9] = {g[4], (T)0.5 * g[5], (T)0.5 * g[6], (T)0.5 * g[5], g[7], (T)0.5 * g[8],
                             (T)0.5 * g[6], (T)0.5 * g[8], g[9]};
            // dE/dlocal multipoles
            T gl[10];
            gl[0] = g[0];
#pragma unroll
            for (int i = 0; i < 3; ++i) gl[1 + i] = R[3 * i] * g[1] + R[3 * i + 1] * g[2] + R[3 * i + 2] * g[3];
            T RG[9];   // RG[i][b] = sum_a R[i][a] Gs[a][b]
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int b = 0; b < 3; ++b) RG[3 * i + b] = R[3 * i] * Gs[b] + R[3 * i + 1] * Gs[3 + b] + R[3 * i + 2] * Gs[6 + b];
            auto gt = [&](int i, int j) { return RG[3 * i] * R[3 * j] + RG[3 * i + 1] * R[3 * j + 1] + RG[3 * i + 2] * R[3 * j + 2]; };
            gl[4] = gt(0, 0); gl[5] = 2 * gt(0, 1); gl[6] = 2 * gt(0, 2); gl[7] = gt(1, 1); gl[8] = 2 * gt(1, 2); gl[9] = gt(2, 2);
            cart_grad_to_harm(gl, dQl + (size_t)a * nh, nh);
            // dE/dR[i][a] = mu_l[i] g_mu[a] + 2 sum_j Tl[i][j] (R Gs)[j][a]
            const T Tl[9] = {ml[4], ml[5], ml[6], ml[5], ml[7], ml[8], ml[6], ml[8], ml[9]};
            V3<T> gr[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                T v[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    v[c] = ml[1 + i] * g[1 + c] + 2 * (Tl[3 * i] * RG[c] + Tl[3 * i + 1] * RG[3 + c] + Tl[3 * i + 2] * RG[6 + c]);
                gr[i] = mk(v[0], v[1], v[2]);
            }
            // reverse of: vy = vz x vxp ; vxp = unit(vx - vz s), s = vx.vz
            V3<T> gz = gr[2] + cross(f.vxp, gr[1]);
            V3<T> gxp = gr[0] + cross(gr[1], f.vz);
            V3<T> gw = unit_bwd(gxp, f.vxp, f.nw);
            T gwz = dot(gw, f.vz);
            V3<T> gvx = gw - f.vz * gwz;
            gz = gz - gw * f.s - f.vx * gwz;
            V3<T> g_vz0 = gz, g_vx0 = gvx, g_vy0 = mk((T)0, (T)0, (T)0);
            if (type == BISECTOR) { V3<T> gs = unit_bwd(gz, f.vz, f.nzb); g_vz0 = gs; g_vx0 = gvx + gs; }
            if (type == ZBISECT) { V3<T> gs = unit_bwd(gvx, f.vx, f.nxb); g_vx0 = gs; g_vy0 = gs; }
            if (type == THREEFOLD) { V3<T> gs = unit_bwd(gz, f.vz, f.nzb); g_vz0 = gs; g_vx0 = gvx + gs; g_vy0 = gs; }
            V3<T> self = mk((T)0, (T)0, (T)0);
            auto push = [&](V3<T> gv, V3<T> v0, T nrm, int anchor, const T (&sh)[3]) {
                V3<T> gd = unit_bwd(gv, v0, nrm);
                atomicAdd(dpos + 3 * anchor, gd.x); atomicAdd(dpos + 3 * anchor + 1, gd.y); atomicAdd(dpos + 3 * anchor + 2, gd.z);
                self = self - gd;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    dbox[3 * p] -= (double)(sh[p] * gd.x); dbox[3 * p + 1] -= (double)(sh[p] * gd.y); dbox[3 * p + 2] -= (double)(sh[p] * gd.z);
                }
            };
            push(g_vz0, f.vz0, f.nz, f.az, f.shz);
            if (type != ZONLY) push(g_vx0, f.vx0, f.nx, f.ax, f.shx);
            if (type == ZBISECT || type == THREEFOLD) push(g_vy0, f.vy0, f.ny, f.ay, f.shy);
            atomicAdd(dpos + 3 * a, self.x); atomicAdd(dpos + 3 * a + 1, self.y); atomicAdd(dpos + 3 * a + 2, self.z);
        }
    }
    if (want_box) block_accumulate<9>(dbox, red, scalars + ADMP_S_DBOX);
}

// rot_global2local / rot_local2global with caller-supplied frames (admp/multipole.py:92-201):
// harmonic -> Cartesian, R (or R^T) applied as a Cartesian rotation, back to harmonic. Equal to
// the reference's polynomial 5x5 matrix for orthonormal frames.
template <typename T>
__global__ void __launch_bounds__(128)
rotate_kernel(int64_t n, int lmax, int to_local, const T* __restrict__ Q, const T* __restrict__ Fr, T* __restrict__ out) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int nh = (lmax + 1) * (lmax + 1);
    T m[10], r[10], R[9];
    harm_to_cart(Q + a * nh, lmax, m);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Fr[a * 9 + k];
    if (!to_local) {   // use the transposed frame
        T t;
        t = R[1]; R[1] = R[3]; R[3] = t; t = R[2]; R[2] = R[6]; R[6] = t; t = R[5]; R[5] = R[7]; R[7] = t;
    }
    r[0] = m[0];
#pragma unroll
    for (int i = 0; i < 3; ++i) r[1 + i] = R[3 * i] * m[1] + R[3 * i + 1] * m[2] + R[3 * i + 2] * m[3];
    const T Tm[9] = {m[4], m[5], m[6], m[5], m[7], m[8], m[6], m[8], m[9]};
    T RT[9];   // RT[i][b] = sum_a R[i][a] T[a][b]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int b = 0; b < 3; ++b) RT[3 * i + b] = R[3 * i] * Tm[b] + R[3 * i + 1] * Tm[3 + b] + R[3 * i + 2] * Tm[6 + b];
    auto tl = [&](int i, int j) { return RT[3 * i] * R[3 * j] + RT[3 * i + 1] * R[3 * j + 1] + RT[3 * i + 2] * R[3 * j + 2]; };
    r[4] = tl(0, 0); r[5] = tl(0, 1); r[6] = tl(0, 2); r[7] = tl(1, 1); r[8] = tl(1, 2); r[9] = tl(2, 2);
    cart_to_harm(r, out + a * nh, nh);
}

template <typename T>
void launch_rotate(cudaStream_t st, int64_t n, int lmax, int to_local, const void* Q, const void* Fr, void* out) {
    if (n <= 0) return;
    rotate_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, lmax, to_local, (const T*)Q, (const T*)Fr, (T*)out);
}
template void launch_rotate<double>(cudaStream_t, int64_t, int, int, const void*, const void*, void*);
template void launch_rotate<float>(cudaStream_t, int64_t, int, int, const void*, const void*, void*);

template <typename T>
void launch_frames_fwd(cudaStream_t st, int n, int lmax, const BoxInfo* B, const void* pos, const int32_t* atype,
                       const int32_t* ai, const void* Ql, void* M, void* Qg, void* Fr) {
    if (n <= 0) return;
    frames_fwd_kernel<T><<<(n + 127) / 128, 128, 0, st>>>(n, lmax, B, (const T*)pos, atype, ai, (const T*)Ql, (T*)M, (T*)Qg, (T*)Fr);
}
template <typename T>
void launch_frames_bwd(cudaStream_t st, int n, int lmax, const BoxInfo* B, const void* pos, const int32_t* atype,
                       const int32_t* ai, const void* Ql, const void* G, void* dQl, void* dpos, double* scalars, int want_box,
                       int first) {
    if (n - first <= 0) return;
    frames_bwd_kernel<T><<<(n - first + 127) / 128, 128, 0, st>>>(first, n, lmax, B, (const T*)pos, atype, ai, (const T*)Ql, (const T*)G,
                                                                  (T*)dQl, (T*)dpos, scalars, want_box);
}
template void launch_frames_fwd<double>(cudaStream_t, int, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const void*, void*, void*, void*);
template void launch_frames_fwd<float>(cudaStream_t, int, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const void*, void*, void*, void*);
template void launch_frames_bwd<double>(cudaStream_t, int, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const void*, const void*, void*, void*, double*, int, int);
template void launch_frames_bwd<float>(cudaStream_t, int, int, const BoxInfo*, const void*, const int32_t*, const int32_t*, const void*, const void*, void*, void*, double*, int, int);

}  // namespace admp
