// C ABI of libadmp_b200.so: context, FFT set-up (hand-written passes; cuFFT plans are created lazily, only when the
// library backend is selected or the mesh sizes are outside the hand-written family), stage entry points and the fused
// evaluation (energy_pme + optimize_Uind + all adjoints) - see include/admp_b200.h.
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

using namespace admp;

static thread_local std::string g_err;

static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKFFT(call)                                                                                \
    do {                                                                                           \
        cufftResult r_ = (call);                                                                   \
        if (r_ != CUFFT_SUCCESS) return fail("%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r_); \
    } while (0)
#define CKLAUNCH() CK(cudaGetLastError())

struct admp_ctx {
    int device = 0, dtype = ADMP_F64, n_sm = 148;
    size_t w = 8;
    double kappa = 0.0;
    int K[3] = {0, 0, 0};
    int lmax = 2;
    int kvec_ref = 0;               // admp_ctx_set_kvec_order
    int n_atoms = 0;
    // topology
    int32_t *axis_type = nullptr, *axis_idx = nullptr, *cov_off = nullptr, *cov_idx = nullptr;
    int8_t* cov_nb = nullptr;
    // cell
    BoxInfo* box = nullptr;
    // reciprocal space
    void *mesh = nullptr, *spec = nullptr, *fftwork = nullptr;
    void* phi = nullptr;            // big meshes: the SCF body's inverse transform writes the potential here (see scf_body)
    const void* phi_cur = nullptr;  // where the potential of the last reciprocal pass lives (mesh or phi)
    cudaEvent_t ev_zfwd = nullptr;
    size_t mesh_bytes = 0, spec_bytes = 0, fftwork_bytes = 0;
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool plans = false;
    double* bt[3] = {nullptr, nullptr, nullptr};
    double *ek = nullptr, *k2 = nullptr;     // separable influence-function tables (per evaluation)
    int* ortho = nullptr;
    ConvTables tb = {};
    Fft3d* fft = nullptr;                    // hand-written FFT (nullptr: sizes unsupported -> cuFFT)
    bool use_custom_fft = false;
    std::string fft_note;
    // per-atom workspaces and staged inputs of admp_pme_eval
    void *M = nullptr, *G = nullptr, *Fscf = nullptr, *rec = nullptr;
    int phi_zld = 0;                // reals per line of phi_cur (0: K3)
    int in_flight = 1;              // admp_ctx_set_in_flight: evaluations the caller keeps in flight on other contexts / streams
    bool inplace_ok = false;        // the real mesh of the fused evaluations may live in the spectrum buffer (set_pme / ADMP_MESH_INPLACE)
    double* cg = nullptr;           // conjugate-gradient work vectors [U | r | p] + rz (allocated on first use, ADMP_SCF_CG)
    void *s_pos = nullptr, *s_U = nullptr, *s_pol = nullptr, *s_th = nullptr, *s_mS = nullptr, *s_pS = nullptr, *s_box = nullptr;
    int32_t* s_pairs = nullptr;
    int8_t* s_sidx = nullptr;       // scale index per staged pair row (-1: row not evaluated)
    int64_t pairs_cap = 0;
    double* scal = nullptr;
    int32_t* state = nullptr;
    int32_t* h_state = nullptr;     // pinned mirror for the host-synchronised loop
    // SCF graph cache
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t cap_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int g_maxiter = -1;
    double g_thresh = -1.0;
    uint32_t g_flags = 0;
    bool graph_failed = false;
    // x-slab decomposition over the GPUs of one NVLink domain (admp_ctx_set_peers)
    int peer_rank = 0, peer_n = 0;
    PeerTab mesh_peers = {}, spec_peers = {};
    SlabAux slab_aux = {};
    // neighbour list
    NbWork nb = {};
    // cluster pair tiles (pair_cluster.cu)
    ClusterWork cw = {};
    int cluster_force = 0;          // ADMP_PAIR_CLUSTER: 0 auto, 1 always (when the row order allows), -1 never
    // brick-staged spread (spread_brick.cu): atom bins per evaluation
    BrickWork bw = {};
    int spread_mode = 0;            // admp_ctx_set_spread: 0 one warp per atom (+ zero-fill; the faster one, see spread_brick.cu), 1 bricks
    size_t ws_bytes = 0;
};

extern "C" const char* admp_last_error(void) { return g_err.c_str(); }
extern "C" int admp_version(void) { return 200; }
static void drop_graph(admp_ctx* c);
extern "C" int admp_ctx_pair_cluster_active(admp_ctx* c) {
    if (!c || !c->cw.state) return -1;
    int32_t h[4] = {0, 0, 0, 0};
    cudaSetDevice(c->device);
    if (cudaMemcpy(h, c->cw.state, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;   // synchronises: tests / tools only
    return h[2];
}
extern "C" int admp_ctx_set_in_flight(admp_ctx* c, int n) {
    if (!c) return fail("admp_ctx_set_in_flight: null context");
    if (n < 1) return fail("admp_ctx_set_in_flight: n = %d", n);
    if ((n > 1) != (c->in_flight > 1)) drop_graph(c);          // the captured SCF body holds the gather variant
    c->in_flight = n;
    return 0;
}
extern "C" int admp_ctx_set_pair_cluster(admp_ctx* c, int force, int min_rows_per_cluster) {
    if (!c) return fail("admp_ctx_set_pair_cluster: null context");
    c->cluster_force = force > 0 ? 1 : (force < 0 ? -1 : 0);
    if (min_rows_per_cluster > 0) c->cw.min_rows_per_cluster = min_rows_per_cluster;
    drop_graph(c);
    return 0;
}
extern "C" int admp_ctx_set_kvec_order(admp_ctx* c, int reference) {
    if (!c) return fail("admp_ctx_set_kvec_order: null context");
    c->kvec_ref = reference ? 1 : 0;
    return 0;
}

static void drop_graph(admp_ctx* c) {
    if (c->gexec) cudaGraphExecDestroy(c->gexec);
    if (c->graph) cudaGraphDestroy(c->graph);
    c->gexec = nullptr;
    c->graph = nullptr;
}

template <typename P> static void dfree(P*& p) {
    if (p) cudaFree((void*)p);
    p = nullptr;
}

extern "C" int admp_ctx_create(admp_ctx** out, int device, int dtype) {
    if (!out) return fail("admp_ctx_create: null out");
    if (dtype != ADMP_F64 && dtype != ADMP_F32) return fail("admp_ctx_create: bad dtype %d", dtype);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("admp_ctx_create: device %d not present (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    admp_ctx* c = new admp_ctx();
    c->device = device;
    c->dtype = dtype;
    c->w = dtype == ADMP_F64 ? 8 : 4;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    CK(cudaMalloc(&c->box, sizeof(BoxInfo)));
    CK(cudaMalloc(&c->scal, sizeof(double) * ADMP_S_COUNT));
    CK(cudaMalloc(&c->state, sizeof(int32_t) * 8));
    CK(cudaMalloc(&c->cw.state, sizeof(int32_t) * 8));
    CK(cudaMemset(c->cw.state, 0, sizeof(int32_t) * 8));
    c->cw.min_rows_per_cluster = 96;
    if (const char* e = getenv("ADMP_PAIR_CLUSTER")) c->cluster_force = atoi(e) > 0 ? 1 : (atoi(e) < 0 || e[0] == '0' ? -1 : 0);
    if (const char* e = getenv("ADMP_PAIR_CLUSTER_MINROWS")) c->cw.min_rows_per_cluster = atoi(e);
    if (const char* e = getenv("ADMP_SPREAD")) c->spread_mode = strcmp(e, "bricks") == 0 ? 1 : 0;
    CK(cudaMalloc(&c->s_mS, 8 * 8));
    CK(cudaMalloc(&c->s_pS, 8 * 8));
    CK(cudaMalloc(&c->s_box, 9 * 8));
    CK(cudaMallocHost(&c->h_state, sizeof(int32_t) * 8));
    CK(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_zfwd, cudaEventDisableTiming));
    c->ws_bytes = sizeof(BoxInfo) + sizeof(double) * ADMP_S_COUNT + 256;
    *out = c;
    return 0;
}

static void free_recip(admp_ctx* c) {
    if (c->plans) { cufftDestroy(c->plan_fwd); cufftDestroy(c->plan_inv); c->plans = false; }
    c->ws_bytes -= c->mesh_bytes + c->spec_bytes + c->fftwork_bytes + (c->phi ? c->mesh_bytes : 0);
    dfree(c->mesh); dfree(c->spec); dfree(c->fftwork); dfree(c->phi);
    c->mesh_bytes = c->spec_bytes = c->fftwork_bytes = 0;
    for (int d = 0; d < 3; ++d) dfree(c->bt[d]);
    dfree(c->ek); dfree(c->k2); dfree(c->ortho);
    if (c->fft) { fft3d_destroy(c->fft); c->fft = nullptr; }
    c->use_custom_fft = false;
    c->peer_n = 0;
    c->ws_bytes -= brick_bytes(c->bw);
    brick_free(c->bw);
}

// cuFFT plans + work area of the library backend: created on first use (the hand-written passes need none of it;
// at 616x1232x1232 the work area alone is 7.5 GB and the two plans take seconds to build)
static int ensure_cufft(admp_ctx* c) {
    if (c->plans) return 0;
    if (!c->mesh) return fail("admp_ctx_set_pme has not been called");
    size_t ws1 = 0, ws2 = 0;
    CKFFT(cufftCreate(&c->plan_fwd));
    CKFFT(cufftCreate(&c->plan_inv));
    c->plans = true;
    CKFFT(cufftSetAutoAllocation(c->plan_fwd, 0));
    CKFFT(cufftSetAutoAllocation(c->plan_inv, 0));
    CKFFT(cufftMakePlan3d(c->plan_fwd, c->K[0], c->K[1], c->K[2], c->dtype == ADMP_F64 ? CUFFT_D2Z : CUFFT_R2C, &ws1));
    CKFFT(cufftMakePlan3d(c->plan_inv, c->K[0], c->K[1], c->K[2], c->dtype == ADMP_F64 ? CUFFT_Z2D : CUFFT_C2R, &ws2));
    c->fftwork_bytes = ws1 > ws2 ? ws1 : ws2;
    if (c->fftwork_bytes == 0) c->fftwork_bytes = 256;
    CK(cudaMalloc(&c->fftwork, c->fftwork_bytes));        // one work area shared by both directions
    CKFFT(cufftSetWorkArea(c->plan_fwd, c->fftwork));
    CKFFT(cufftSetWorkArea(c->plan_inv, c->fftwork));
    c->ws_bytes += c->fftwork_bytes;
    return 0;
}

static void free_atoms(admp_ctx* c) {
    dfree(c->M); dfree(c->G); dfree(c->Fscf); dfree(c->rec); dfree(c->cg); dfree(c->s_pos); dfree(c->s_U); dfree(c->s_pol); dfree(c->s_th);
    dfree(c->axis_type); dfree(c->axis_idx); dfree(c->cov_off); dfree(c->cov_idx); dfree(c->cov_nb);
    dfree(c->cw.cl_of); dfree(c->cw.cl_first); dfree(c->cw.cl_size); dfree(c->cw.row_start); dfree(c->cw.cl_extra);
    c->cw.n_clusters = 0;
    c->ws_bytes -= brick_bytes(c->bw);
    brick_free(c->bw);
}

extern "C" int admp_ctx_destroy(admp_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    drop_graph(c);
    free_recip(c);
    free_atoms(c);
    dfree(c->s_pairs); dfree(c->s_sidx); dfree(c->cw.ent_i); dfree(c->cw.ent_m); dfree(c->cw.state); dfree(c->box); dfree(c->scal); dfree(c->state); dfree(c->s_mS); dfree(c->s_pS); dfree(c->s_box);
    dfree(c->nb.cell_of); dfree(c->nb.cell_count); dfree(c->nb.cell_start); dfree(c->nb.sorted); dfree(c->nb.nbr_count); dfree(c->nb.nbr_start); dfree(c->nb.geom); dfree(c->nb.scan_tmp);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_zfwd) cudaEventDestroy(c->ev_zfwd);
    if (c->slab_aux.ready) {
        for (int s = 0; s < SLAB_STREAMS; ++s) cudaStreamDestroy(c->slab_aux.copy_stream[s]);
        cudaEventDestroy(c->slab_aux.fork);
        for (int k = 0; k < SLAB_CHUNKS; ++k)
            for (int s = 0; s < SLAB_STREAMS; ++s) cudaEventDestroy(c->slab_aux.chunk[k][s]);
        for (int k = 0; k < SLAB_CHUNKS; ++k) cudaEventDestroy(c->slab_aux.done[k]);
        cudaEventDestroy(c->slab_aux.pushed);
    }
    delete c;
    return 0;
}

extern "C" int64_t admp_ctx_workspace_bytes(const admp_ctx* c) { return c ? (int64_t)c->ws_bytes : 0; }
extern "C" int admp_ctx_scf_graph_active(const admp_ctx* c) { return (c && c->gexec && !c->graph_failed) ? 1 : 0; }

// 1/theta_k^2 per dimension (admp/recip.py:400-408): theta = sum_{m=-2..2} M6(m+3) cos(2 pi m k / K)
static std::vector<double> theta_inv2(int K, int count) {
    const double M6[5] = {1.0 / 120, 26.0 / 120, 66.0 / 120, 26.0 / 120, 1.0 / 120};   // M6(1..5)
    std::vector<double> out(count);
    for (int i = 0; i < count; ++i) {
        const int k = (2 * i < K) ? i : i - K;
        double th = 0.0;
        for (int m = -2; m <= 2; ++m) th += M6[m + 2] * std::cos(2.0 * M_PI * m * (double)k / (double)K);
        out[i] = 1.0 / (th * th);
    }
    return out;
}

// atom bins of the brick-staged spread: need both the mesh and the atom count (whichever of set_pme / set_topology comes second)
static int ensure_bricks(admp_ctx* c) {
    if (c->bw.ready || !c->mesh || c->n_atoms <= 0) return 0;
    if (c->spread_mode == 0) return 0;                 // allocated only when the brick variant is asked for
    CK(brick_alloc(c->bw, c->n_atoms, c->K, c->n_sm, c->w));
    c->ws_bytes += brick_bytes(c->bw);
    return 0;
}
static inline bool use_bricks(const admp_ctx* c) { return c->bw.ready && c->spread_mode == 1; }

extern "C" int admp_ctx_set_spread(admp_ctx* c, int bricks) {
    if (!c) return fail("admp_ctx_set_spread: null context");
    CK(cudaSetDevice(c->device));
    c->spread_mode = bricks ? 1 : 0;
    drop_graph(c);
    return ensure_bricks(c);
}
extern "C" int admp_ctx_spread_bricks(const admp_ctx* c) { return (c && use_bricks(c)) ? c->bw.geom.bz : 0; }

extern "C" int admp_ctx_set_pme(admp_ctx* c, double kappa, int K1, int K2, int K3, int lmax) {
    if (!c) return fail("admp_ctx_set_pme: null ctx");
    if (lmax < 0 || lmax > 2) return fail("l > 2 (beyond quadrupole) not supported");   // admp/multipole.py:111
    if (K1 < 6 || K2 < 6 || K3 < 6) return fail("admp_ctx_set_pme: mesh %dx%dx%d smaller than the order-6 stencil", K1, K2, K3);
    if (!(kappa > 0.0)) return fail("admp_ctx_set_pme: kappa must be positive");
    CK(cudaSetDevice(c->device));
    c->kappa = kappa;
    c->lmax = lmax;
    drop_graph(c);
    if (c->mesh && c->K[0] == K1 && c->K[1] == K2 && c->K[2] == K3) return 0;
    free_recip(c);
    c->K[0] = K1; c->K[1] = K2; c->K[2] = K3;
    const size_t G = (size_t)K1 * K2 * K3, Gh = (size_t)K1 * K2 * (K3 / 2 + 1);
    c->mesh_bytes = G * c->w;
    c->spec_bytes = Gh * 2 * c->w;
    CK(cudaMalloc(&c->mesh, c->mesh_bytes));
    CK(cudaMalloc(&c->spec, c->spec_bytes));
    const int cnt[3] = {K1, K2, K3 / 2 + 1};
    for (int d = 0; d < 3; ++d) {
        std::vector<double> t = theta_inv2(c->K[d], cnt[d]);
        CK(cudaMalloc(&c->bt[d], sizeof(double) * cnt[d]));
        CK(cudaMemcpy(c->bt[d], t.data(), sizeof(double) * cnt[d], cudaMemcpyHostToDevice));
    }
    const int ntab = cnt[0] + cnt[1] + cnt[2];
    CK(cudaMalloc(&c->ek, sizeof(double) * ntab));
    CK(cudaMalloc(&c->k2, sizeof(double) * ntab));
    CK(cudaMalloc(&c->ortho, sizeof(int)));
    int off = 0;
    for (int d = 0; d < 3; ++d) {
        c->tb.ek[d] = c->ek + off; c->tb.k2[d] = c->k2 + off; c->tb.bt[d] = c->bt[d];
        off += cnt[d];
    }
    c->tb.ortho = c->ortho;
    // hand-written FFT fused with the convolution whenever the mesh sizes allow it
    const char* why = "";
    const char* env = getenv("ADMP_FFT");
    c->fft = fft3d_create(K1, K2, K3, c->dtype, &why);
    c->use_custom_fft = (c->fft != nullptr) && !(env && strcmp(env, "cufft") == 0);
    c->fft_note = c->fft ? "" : why;
    {
        const char* e3 = getenv("ADMP_MESH_INPLACE");
        c->inplace_ok = c->fft != nullptr && fft3d_inplace_supported(c->fft) && !(e3 && atoi(e3) == 0);
    }
    cudaGetLastError();
    c->ws_bytes += c->mesh_bytes + c->spec_bytes;
    // optional second real buffer for the SCF body (ADMP_TWO_MESH=1): measured and left off - on the L2-resident 154^3 mesh
    // the larger working set costs 8 %, on 308x616x616 the overlapped zero-fill gains 1 % (97.2 vs 98.2 ms per evaluation:
    // the memset kernel waits for SM slots behind the persistent X-pass blocks), not worth one more mesh of memory
    {
        const char* e2 = getenv("ADMP_TWO_MESH");
        const bool want = e2 ? atoi(e2) > 0 : false;
        if (want && c->use_custom_fft) {
            if (cudaMalloc(&c->phi, c->mesh_bytes) == cudaSuccess) c->ws_bytes += c->mesh_bytes;
            else { c->phi = nullptr; cudaGetLastError(); }
        }
    }
    if (!c->use_custom_fft && ensure_cufft(c)) return 1;
    return ensure_bricks(c);
}

/* 1: hand-written fused FFT, 0: cuFFT + separate convolution kernel */
extern "C" int admp_ctx_fft_backend(const admp_ctx* c) { return (c && c->use_custom_fft) ? 1 : 0; }
extern "C" int admp_ctx_set_fft_backend(admp_ctx* c, int custom) {
    if (!c) return fail("null ctx");
    if (custom && !c->fft) return fail("hand-written FFT unavailable for this mesh: %s", c->fft_note.c_str());
    if (!custom && ensure_cufft(c)) return 1;           // plan creation allocates: never inside a stream capture
    c->use_custom_fft = custom != 0;
    drop_graph(c);
    return 0;
}

extern "C" int admp_ctx_set_topology(admp_ctx* c, int n, const int32_t* axis_type, const int32_t* axis_indices,
                                     const int32_t* cov_offsets, const int32_t* cov_index, const int8_t* cov_nbonds) {
    if (!c) return fail("admp_ctx_set_topology: null ctx");
    if (n <= 0) return fail("admp_ctx_set_topology: n_atoms = %d", n);
    CK(cudaSetDevice(c->device));
    drop_graph(c);
    free_atoms(c);
    c->n_atoms = n;
    const size_t w = c->w;
    CK(cudaMalloc(&c->M, (size_t)n * 10 * w));
    CK(cudaMalloc(&c->G, (size_t)n * 10 * w));
    CK(cudaMalloc(&c->Fscf, (size_t)n * 3 * w));
    CK(cudaMalloc(&c->rec, (size_t)n * pair_record_bytes((int)w)));
    CK(cudaMalloc(&c->s_pos, (size_t)n * 3 * w));
    CK(cudaMalloc(&c->s_U, (size_t)n * 3 * w));
    CK(cudaMalloc(&c->s_pol, (size_t)n * w));
    CK(cudaMalloc(&c->s_th, (size_t)n * w));
    if (axis_type && axis_indices) {
        CK(cudaMalloc(&c->axis_type, sizeof(int32_t) * n));
        CK(cudaMalloc(&c->axis_idx, sizeof(int32_t) * 3 * n));
        // anchors a site's axis type does not use may be -1 (admp/parser.py); keep reads in range
        std::vector<int32_t> ai(axis_indices, axis_indices + 3 * (size_t)n);
        for (size_t k = 0; k < ai.size(); ++k)
            if (ai[k] < 0 || ai[k] >= n) ai[k] = (int32_t)(k / 3);
        for (int a = 0; a < n; ++a) {
            const int t = axis_type[a];
            if (t < 0 || t > 5) return fail("admp_ctx_set_topology: axis type %d of atom %d out of range", t, a);
            const int need = (t == 5) ? 0 : (t == 4) ? 1 : (t == 2 || t == 3) ? 3 : 2;
            for (int k = 0; k < need; ++k) {
                const int v = axis_indices[3 * a + k];
                if (v < 0 || v >= n || v == a) return fail("admp_ctx_set_topology: atom %d (axis type %d) has invalid anchor %d", a, t, v);
            }
        }
        CK(cudaMemcpy(c->axis_type, axis_type, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->axis_idx, ai.data(), sizeof(int32_t) * 3 * n, cudaMemcpyHostToDevice));
    }
    if (cov_offsets && cov_offsets[n] > 0) {
        if (!cov_index || !cov_nbonds) return fail("admp_ctx_set_topology: covalent CSR incomplete");
        const int nnz = cov_offsets[n];
        CK(cudaMalloc(&c->cov_off, sizeof(int32_t) * (n + 1)));
        CK(cudaMalloc(&c->cov_idx, sizeof(int32_t) * nnz));
        CK(cudaMalloc(&c->cov_nb, nnz));
        CK(cudaMemcpy(c->cov_off, cov_offsets, sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->cov_idx, cov_index, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->cov_nb, cov_nbonds, nnz, cudaMemcpyHostToDevice));
    }
    // j-clusters of the cluster pair kernel: runs of consecutive atoms (at most 4) in which every atom is covalently
    // listed with an earlier member - a water molecule, a methyl group ... (singletons without a covalent map)
    {
        std::vector<int32_t> cl_of(n), cl_first, cl_size;
        int a = 0;
        while (a < n) {
            int sz = 1;
            while (sz < 4 && a + sz < n && cov_offsets && cov_index) {
                const int b = a + sz;
                bool bonded = false;
                for (int k = cov_offsets[b]; k < cov_offsets[b + 1] && !bonded; ++k) bonded = (cov_index[k] >= a && cov_index[k] < b);
                if (!bonded) break;
                ++sz;
            }
            for (int k = 0; k < sz; ++k) cl_of[a + k] = (int32_t)cl_first.size();
            cl_first.push_back(a);
            cl_size.push_back(sz);
            a += sz;
        }
        const int nc = (int)cl_first.size();
        c->cw.n_clusters = nc;
        CK(cudaMalloc(&c->cw.cl_of, sizeof(int32_t) * n));
        CK(cudaMalloc(&c->cw.cl_first, sizeof(int32_t) * nc));
        CK(cudaMalloc(&c->cw.cl_size, sizeof(int32_t) * nc));
        CK(cudaMalloc(&c->cw.cl_extra, sizeof(int32_t) * nc));
        CK(cudaMalloc(&c->cw.row_start, sizeof(int32_t) * ((size_t)n + 2)));
        CK(cudaMemcpy(c->cw.cl_of, cl_of.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->cw.cl_first, cl_first.data(), sizeof(int32_t) * nc, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->cw.cl_size, cl_size.data(), sizeof(int32_t) * nc, cudaMemcpyHostToDevice));
    }
    return ensure_bricks(c);
}

// ------------------------------------------------------------------------------------------ helpers
#define DISPATCH(c, fn, ...)                              \
    do {                                                  \
        if ((c)->dtype == ADMP_F64) fn<double>(__VA_ARGS__); \
        else fn<float>(__VA_ARGS__);                      \
    } while (0)

static int need(admp_ctx* c, bool recip, bool atoms) {
    if (!c) return fail("null ctx");
    if (recip && !c->mesh) return fail("admp_ctx_set_pme has not been called");
    if (atoms && c->n_atoms <= 0) return fail("admp_ctx_set_topology has not been called");
    return 0;
}

static int fft_fwd(admp_ctx* c, cudaStream_t st) {
    if (ensure_cufft(c)) return 1;
    CKFFT(cufftSetStream(c->plan_fwd, st));
    if (c->dtype == ADMP_F64) CKFFT(cufftExecD2Z(c->plan_fwd, (cufftDoubleReal*)c->mesh, (cufftDoubleComplex*)c->spec));
    else CKFFT(cufftExecR2C(c->plan_fwd, (cufftReal*)c->mesh, (cufftComplex*)c->spec));
    return 0;
}
static int fft_inv(admp_ctx* c, cudaStream_t st) {
    if (ensure_cufft(c)) return 1;
    CKFFT(cufftSetStream(c->plan_inv, st));
    if (c->dtype == ADMP_F64) CKFFT(cufftExecZ2D(c->plan_inv, (cufftDoubleComplex*)c->spec, (cufftDoubleReal*)c->mesh));
    else CKFFT(cufftExecC2R(c->plan_inv, (cufftComplex*)c->spec, (cufftReal*)c->mesh));
    return 0;
}

// spread -> FFT -> influence function (+energy) -> inverse FFT; leaves phi = dE/dmesh in c->mesh
static void conv_tables(admp_ctx* c, cudaStream_t st) {
    const int maxK = std::max(c->K[0], std::max(c->K[1], c->K[2]));
    launch_conv_tables(st, c->box, c->kappa, c->bt[0], c->bt[1], c->bt[2], c->ek, c->k2, c->ortho, maxK);
}

// once per evaluation (positions / box changed): bin the atoms by mesh brick
static void spread_prepare(admp_ctx* c, cudaStream_t st, const void* pos) {
    if (use_bricks(c)) DISPATCH(c, launch_brick_sort, st, c->bw, c->box, pos);
}
// the whole spread: mesh = sum of the atoms' stencils (bricks: one coalesced write; otherwise zero-fill + per-atom scatter)
// In-place mesh of the fused evaluations (admp_pme_eval, admp_disp_eval, admp_pme_recip): the real mesh occupies the spectrum
// buffer, line l of K3 reals at the start of spectrum line l (the layout of an in-place real-to-complex transform), so one
// reciprocal round trip touches 16 (K3/2 + 1) K1 K2 bytes instead of twice that - on the 154^3 mesh of the reference's examples four
// evaluations in flight then fit the 126 MB L2. Needs the tile-pipelined Z passes (they hold a whole tile in shared memory before
// they write it); per-atom spread only; off for the stand-alone stage entry points, the slab / atom-block decompositions and the
// two-buffer SCF body, which keep the separate mesh. ADMP_MESH_INPLACE=0 switches it off.
static bool mesh_inplace(const admp_ctx* c) {
    return c->inplace_ok && c->use_custom_fft && c->fft && c->phi == nullptr && !use_bricks(c) && c->spec_peers.n <= 1;
}
static void* mesh_buf(admp_ctx* c) { return mesh_inplace(c) ? c->spec : c->mesh; }
static int mesh_zld(const admp_ctx* c) { return mesh_inplace(c) ? 2 * (c->K[2] / 2 + 1) : 0; }
static size_t mesh_fill_bytes(const admp_ctx* c) { return mesh_inplace(c) ? c->spec_bytes : c->mesh_bytes; }

// fused: part of a self-contained spread -> transform -> gather chain (may use the in-place mesh); the stand-alone stage entry
// points always work on the separate mesh buffer
static int spread_all(admp_ctx* c, cudaStream_t st, const void* pos, const void* M, int cols, int stride, const void* U, bool fused) {
    if (use_bricks(c)) {
        DISPATCH(c, launch_spread_brick, st, c->bw, c->box, pos, M, cols, stride, U, c->mesh);
    } else if (fused) {
        CK(cudaMemsetAsync(mesh_buf(c), 0, mesh_fill_bytes(c), st));
        DISPATCH(c, launch_spread, st, c->n_atoms, c->box, pos, M, cols, stride, U, mesh_buf(c), nullptr, mesh_zld(c));
    } else {
        CK(cudaMemsetAsync(c->mesh, 0, c->mesh_bytes, st));
        DISPATCH(c, launch_spread, st, c->n_atoms, c->box, pos, M, cols, stride, U, c->mesh);
    }
    CKLAUNCH();
    return 0;
}

// `tables`: rebuild the separable influence tables and the atom bins first (needed once per box / positions, i.e. per evaluation)
static int recip_field(admp_ctx* c, cudaStream_t st, const void* pos, const void* M, int cols, int stride, const void* U,
                       int kind, double* scalars, int want_vir, bool tables = true) {
    if (tables) {
        conv_tables(c, st);
        spread_prepare(c, st, pos);
    }
    c->phi_cur = mesh_buf(c);
    c->phi_zld = mesh_zld(c);
    if (spread_all(c, st, pos, M, cols, stride, U, true)) return 1;
    if (c->use_custom_fft) {
        fft3d_convolve_roundtrip(c->fft, st, mesh_buf(c), c->spec, c->box, c->kappa, kind, c->tb, scalars, want_vir, nullptr, nullptr, mesh_zld(c));
        CKLAUNCH();
        return 0;
    }
    if (fft_fwd(c, st)) return 1;
    const size_t nh = (size_t)c->K[0] * c->K[1] * (c->K[2] / 2 + 1);
    DISPATCH(c, launch_convolve, st, c->box, nh, c->n_sm, c->kappa, kind, c->tb, c->spec, scalars, want_vir);
    CKLAUNCH();
    if (fft_inv(c, st)) return 1;
    return 0;
}

// ------------------------------------------------------------------------------------------ stages
extern "C" int admp_frames_fwd(admp_ctx* c, void* stream, const void* pos, const void* box, const void* Ql, void* M, void* Qg,
                               void* frames) {
    if (need(c, false, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    DISPATCH(c, launch_frames_fwd, st, c->n_atoms, c->lmax, c->box, pos, c->axis_type, c->axis_idx, Ql, M, Qg, frames);
    CKLAUNCH();
    return 0;
}

extern "C" int admp_frames_bwd(admp_ctx* c, void* stream, const void* pos, const void* box, const void* Ql, const void* G,
                               void* dQl, void* dpos, double* scalars) {
    if (need(c, false, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    DISPATCH(c, launch_frames_bwd, st, c->n_atoms, c->lmax, c->box, pos, c->axis_type, c->axis_idx, Ql, G, dQl, dpos, scalars, 1);
    CKLAUNCH();
    return 0;
}

extern "C" int admp_rotate(admp_ctx* c, void* stream, int64_t n, int lmax, int to_local, const void* Q, const void* frames, void* out) {
    if (!c) return fail("null ctx");
    if (lmax < 0 || lmax > 2) return fail("l > 2 (beyond quadrupole) not supported");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_rotate, st, n, lmax, to_local, Q, frames, out);
    CKLAUNCH();
    return 0;
}

static int ensure_pairs(admp_ctx* c, int64_t n_rows);
extern "C" int admp_pme_real(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                             const void* M, const void* U, const void* pol, const void* tholes, const void* mScales,
                             const void* pScales, int mode, uint32_t flags, void* dpos, void* G, void* F, void* dpol,
                             void* dtholes, double* scalars) {
    if (need(c, false, true)) return 1;
    if (U && (!pol || !tholes || !pScales)) return fail("admp_pme_real: polarizable call needs pol, tholes and pScales");
    if (mode == 1 && !U) return fail("admp_pme_real: mode 1 (field only) needs U");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    // scale index per row once per call (one byte per row) instead of a covalent-CSR walk inside the pair loop, then the
    // cluster tiles when the list qualifies (pair_cluster.cu); exactly one of the two pair kernels does the work
    if (!(flags & ADMP_REUSE_PAIR_TILES)) {
        if (ensure_pairs(c, n_rows)) return 1;
        launch_pair_scale(st, n_rows, c->n_atoms, pairs, c->cov_off, c->cov_idx, c->cov_nb, c->s_sidx);
        launch_cluster_prepare(st, n_rows, c->n_atoms, c->cw.n_clusters, pairs, c->s_sidx, c->cw, c->cluster_force);
    } else if (n_rows > c->pairs_cap) {
        return fail("admp_pme_real: ADMP_REUSE_PAIR_TILES without a previous call on this pair list");
    }
    if (mode == 1) DISPATCH(c, launch_pair_pack, st, c->n_atoms, pos, M, U, pol, tholes, c->rec);
    DISPATCH(c, launch_pme_pair, st, n_rows, c->n_atoms, c->box, c->kappa, pos, pairs, c->s_sidx, c->cov_off, c->cov_idx, c->cov_nb, M, U, pol,
             tholes, mScales, pScales, mode, flags, dpos, G, F, dpol, dtholes, scalars, c->rec, c->cw.state);
    DISPATCH(c, launch_pme_cluster, st, c->cw.n_clusters, c->box, c->kappa, c->cw, c->rec, U, mScales, pScales, mode, flags, dpos, G, F, dpol,
             dtholes, scalars);
    CKLAUNCH();
    return 0;
}

extern "C" int admp_pme_recip(admp_ctx* c, void* stream, const void* pos, const void* box, const void* M, int M_cols, int M_stride,
                              const void* U, int kind, int mode, uint32_t flags, void* dpos, void* G, int G_stride, void* F,
                              double* scalars) {
    if (need(c, true, true)) return 1;
    if (M_cols != 1 && M_cols != 10) return fail("admp_pme_recip: M_cols must be 1 or 10");
    if (kind != ADMP_CK_COULOMB && kind != ADMP_CK_DISP6 && kind != ADMP_CK_DISP8 && kind != ADMP_CK_DISP10)
        return fail("admp_pme_recip: unknown influence function %d", kind);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    if (recip_field(c, st, pos, M, M_cols, M_stride, U, kind, scalars, (flags & ADMP_WANT_VIRIAL) ? 1 : 0)) return 1;
    if (mode == 1 || (flags & ADMP_WANT_GRAD)) {
        DISPATCH(c, launch_gather, st, c->n_atoms, c->box, pos, M, M_cols, M_stride, U, c->phi_cur, mode, flags, dpos, G, G_stride, F, scalars,
                 nullptr, c->phi_zld, c->in_flight > 1);
        CKLAUNCH();
    }
    return 0;
}

// ---- individual reciprocal stages on the context's mesh (used by stage-level tests and by the
// per-kernel roofline timing of bench.py; admp_pme_recip chains them)
extern "C" int admp_pme_spread(admp_ctx* c, void* stream, const void* pos, const void* box, const void* M, int M_cols, int M_stride,
                               const void* U) {
    if (need(c, true, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    spread_prepare(c, st, pos);
    return spread_all(c, st, pos, M, M_cols, M_stride, U, false);
}
extern "C" int admp_pme_spread_only(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride, const void* U) {
    if (need(c, true, true)) return 1;     // no box set-up, no binning: the spread kernel alone on the state admp_pme_spread left
    if (use_bricks(c)) {                   // (bricks: writes the whole mesh; per-atom path: the bare scatter without the zero-fill)
        DISPATCH(c, launch_spread_brick, (cudaStream_t)stream, c->bw, c->box, pos, M, M_cols, M_stride, U, c->mesh);
    } else {
        DISPATCH(c, launch_spread, (cudaStream_t)stream, c->n_atoms, c->box, pos, M, M_cols, M_stride, U, c->mesh);
    }
    CKLAUNCH();
    return 0;
}
extern "C" int admp_pme_fft(admp_ctx* c, void* stream, int inverse) {
    if (need(c, true, false)) return 1;
    CK(cudaSetDevice(c->device));
    if (c->use_custom_fft) {
        if (inverse) fft3d_inverse(c->fft, (cudaStream_t)stream, c->spec, c->mesh);
        else fft3d_forward(c->fft, (cudaStream_t)stream, c->mesh, c->spec);
        CKLAUNCH();
        return 0;
    }
    return inverse ? fft_inv(c, (cudaStream_t)stream) : fft_fwd(c, (cudaStream_t)stream);
}
/* the fused five-pass round trip on the context's mesh (mesh -> phi), hand-written FFT only */
extern "C" int admp_pme_fft_convolve(admp_ctx* c, void* stream, int kind, uint32_t flags, double* scalars) {
    if (need(c, true, false)) return 1;
    if (!c->fft) return fail("hand-written FFT unavailable for this mesh: %s", c->fft_note.c_str());
    cudaStream_t st = (cudaStream_t)stream;
    const int maxK = std::max(c->K[0], std::max(c->K[1], c->K[2]));
    launch_conv_tables(st, c->box, c->kappa, c->bt[0], c->bt[1], c->bt[2], c->ek, c->k2, c->ortho, maxK);
    fft3d_convolve_roundtrip(c->fft, st, c->mesh, c->spec, c->box, c->kappa, kind, c->tb, scalars, (flags & ADMP_WANT_VIRIAL) ? 1 : 0);
    CKLAUNCH();
    return 0;
}
/* one of the five passes of the fused round trip (0 Z-fwd, 1 Y-fwd, 2 X-fwd*conv*X-inv, 3 Y-inv, 4 Z-inv) */
extern "C" int admp_pme_fft_pass(admp_ctx* c, void* stream, int which, int kind, double* scalars) {
    if (need(c, true, false)) return 1;
    if (!c->fft) return fail("hand-written FFT unavailable for this mesh: %s", c->fft_note.c_str());
    if (which < 0 || which > 4) return fail("admp_pme_fft_pass: pass index %d", which);
    fft3d_single_pass(c->fft, (cudaStream_t)stream, which, c->mesh, c->spec, c->box, c->kappa, kind, c->tb, scalars);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_pme_convolve(admp_ctx* c, void* stream, int kind, uint32_t flags, double* scalars) {
    if (need(c, true, false)) return 1;
    const size_t nh = (size_t)c->K[0] * c->K[1] * (c->K[2] / 2 + 1);
    const int maxK = std::max(c->K[0], std::max(c->K[1], c->K[2]));
    launch_conv_tables((cudaStream_t)stream, c->box, c->kappa, c->bt[0], c->bt[1], c->bt[2], c->ek, c->k2, c->ortho, maxK);
    DISPATCH(c, launch_convolve, (cudaStream_t)stream, c->box, nh, c->n_sm, c->kappa, kind, c->tb, c->spec, scalars,
             (flags & ADMP_WANT_VIRIAL) ? 1 : 0);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_pme_gather(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride, const void* U,
                               int mode, uint32_t flags, void* dpos, void* G, int G_stride, void* F, double* scalars) {
    if (need(c, true, true)) return 1;
    DISPATCH(c, launch_gather, (cudaStream_t)stream, c->n_atoms, c->box, pos, M, M_cols, M_stride, U, c->mesh, mode, flags, dpos, G,
             G_stride, F, scalars);
    CKLAUNCH();
    return 0;
}
/* which: 0 = real mesh (K1*K2*K3 reals), 1 = half spectrum (K1*K2*(K3/2+1) complex) */
extern "C" void* admp_ctx_buffer(admp_ctx* c, int which) { return !c ? nullptr : (which == 0 ? c->mesh : c->spec); }
/* copy between a caller device buffer and the context's mesh (which = 0) / spectrum (which = 1) */
extern "C" int admp_ctx_buffer_io(admp_ctx* c, void* stream, int which, void* user, int64_t nbytes, int to_ctx) {
    if (need(c, true, false)) return 1;
    void* buf = which == 0 ? c->mesh : c->spec;
    const size_t cap = which == 0 ? c->mesh_bytes : c->spec_bytes;
    if (nbytes < 0 || (size_t)nbytes > cap) return fail("admp_ctx_buffer_io: %lld bytes exceed the buffer (%zu)", (long long)nbytes, cap);
    CK(cudaMemcpyAsync(to_ctx ? buf : user, to_ctx ? user : buf, (size_t)nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int admp_pme_self(admp_ctx* c, void* stream, const void* M, const void* U, const void* pol, uint32_t flags, void* G,
                             void* F, void* dpol, double* scalars) {
    if (need(c, false, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_self, st, c->n_atoms, c->kappa, M, U, pol, flags, G, F, dpol, scalars);
    CKLAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------ atom-range stages
// Building blocks of the multi-GPU atom-block decomposition (admp_b200/parallel.py): every rank holds
// replicated positions / multipoles, works on atoms [first, first+count) and on its slice of the pair
// rows, and the caller reduces mesh / field / gradient arrays across ranks between the stages.
template <typename P> static P* shift(P* p, size_t elems, size_t w) {
    return p ? (P*)((char*)p + elems * w) : nullptr;
}
extern "C" int admp_set_box(admp_ctx* c, void* stream, const void* box) {
    if (need(c, true, false)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    conv_tables(c, st);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_mesh_zero(admp_ctx* c, void* stream) {
    if (need(c, true, false)) return 1;
    CK(cudaMemsetAsync(c->mesh, 0, c->mesh_bytes, (cudaStream_t)stream));
    return 0;
}
extern "C" int admp_pme_spread_range(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride,
                                     const void* U, int first, int count) {
    if (need(c, true, true)) return 1;
    if (first < 0 || count < 0 || first + count > c->n_atoms) return fail("atom range [%d,%d) outside [0,%d)", first, first + count, c->n_atoms);
    const size_t w = c->w;
    DISPATCH(c, launch_spread, (cudaStream_t)stream, count, c->box, shift(pos, (size_t)first * 3, w), shift(M, (size_t)first * M_stride, w),
             M_cols, M_stride, shift(U, (size_t)first * 3, w), c->mesh);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_pme_gather_range(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride, const void* U,
                                     int mode, uint32_t flags, void* dpos, void* G, int G_stride, void* F, double* scalars, int first,
                                     int count) {
    if (need(c, true, true)) return 1;
    if (first < 0 || count < 0 || first + count > c->n_atoms) return fail("atom range [%d,%d) outside [0,%d)", first, first + count, c->n_atoms);
    const size_t w = c->w;
    DISPATCH(c, launch_gather, (cudaStream_t)stream, count, c->box, shift(pos, (size_t)first * 3, w), shift(M, (size_t)first * M_stride, w),
             M_cols, M_stride, shift(U, (size_t)first * 3, w), c->mesh, mode, flags, shift(dpos, (size_t)first * 3, w),
             shift(G, (size_t)first * G_stride, w), G_stride, shift(F, (size_t)first * 3, w), scalars);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_pme_self_range(admp_ctx* c, void* stream, const void* M, const void* U, const void* pol, uint32_t flags, void* G,
                                   void* F, void* dpol, double* scalars, int first, int count) {
    if (need(c, false, true)) return 1;
    if (first < 0 || count < 0 || first + count > c->n_atoms) return fail("atom range [%d,%d) outside [0,%d)", first, first + count, c->n_atoms);
    const size_t w = c->w;
    DISPATCH(c, launch_self, (cudaStream_t)stream, count, c->kappa, shift(M, (size_t)first * 10, w), shift(U, (size_t)first * 3, w),
             shift(pol, (size_t)first, w), flags, shift(G, (size_t)first * 10, w), shift(F, (size_t)first * 3, w), shift(dpol, (size_t)first, w),
             scalars);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_frames_bwd_range(admp_ctx* c, void* stream, const void* pos, const void* Ql, const void* G, void* dQl, void* dpos,
                                     double* scalars, int first, int count) {
    if (need(c, false, true)) return 1;
    if (first < 0 || count < 0 || first + count > c->n_atoms) return fail("atom range [%d,%d) outside [0,%d)", first, first + count, c->n_atoms);
    DISPATCH(c, launch_frames_bwd, (cudaStream_t)stream, first + count, c->lmax, c->box, pos, c->axis_type, c->axis_idx, Ql, G, dQl, dpos,
             scalars, 1, first);
    CKLAUNCH();
    return 0;
}
// ------------------------------------------------------------------------------------------ x-slab stages
// Multi-GPU reciprocal space without a replicated mesh: rank r owns the x planes [r*K1/n, (r+1)*K1/n) of the
// real mesh and of the half spectrum (every rank keeps full-size buffers so the element offsets are the
// single-GPU ones; only the own planes are ever touched locally). Spread and gather address the owner of each
// stencil plane through the peer table (NVLink peer atomics / loads for stencils that cross a slab boundary);
// the Z and Y passes run on the own planes; the fused X pass transforms this rank's share of the (y, kz)
// columns straight out of / into all ranks' planes. The caller (admp_b200/parallel.py) orders the stages with
// stream-ordered cross-rank barriers.
extern "C" int admp_ipc_export(const void* devptr, void* handle64) {
    if (!devptr || !handle64) return fail("admp_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(devptr)));
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int admp_ipc_open(const void* handle64, void** out) {
    if (!handle64 || !out) return fail("admp_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int admp_ipc_close(void* devptr) {
    if (devptr) CK(cudaIpcCloseMemHandle(devptr));
    return 0;
}
extern "C" int admp_ctx_set_peers(admp_ctx* c, int rank, int nranks, void* const* mesh_ptrs, void* const* spec_ptrs) {
    if (need(c, true, false)) return 1;
    if (nranks == 0) { c->peer_n = 0; return 0; }
    if (nranks < 1 || nranks > ADMP_MAX_PEERS) return fail("admp_ctx_set_peers: %d ranks (1..%d supported)", nranks, ADMP_MAX_PEERS);
    if (rank < 0 || rank >= nranks) return fail("admp_ctx_set_peers: rank %d outside [0,%d)", rank, nranks);
    if (c->K[0] < nranks) return fail("admp_ctx_set_peers: K1 = %d is smaller than %d ranks", c->K[0], nranks);
    if (!c->fft || !fft3d_slab_supported(c->fft))
        return fail("admp_ctx_set_peers: the x-slab passes need the register-blocked FFT kernels (mesh family 154*2^n); %s", c->fft_note.c_str());
    if (!mesh_ptrs || !spec_ptrs) return fail("admp_ctx_set_peers: null pointer tables");
    if (mesh_ptrs[rank] != c->mesh || spec_ptrs[rank] != c->spec) return fail("admp_ctx_set_peers: entry %d must be this context's own buffers", rank);
    for (int r = 0; r < ADMP_MAX_PEERS; ++r) {
        c->mesh_peers.base[r] = r < nranks ? mesh_ptrs[r] : nullptr;
        c->spec_peers.base[r] = r < nranks ? spec_ptrs[r] : nullptr;
        if (r < nranks && (!mesh_ptrs[r] || !spec_ptrs[r])) return fail("admp_ctx_set_peers: null buffer for rank %d", r);
    }
    if (!c->slab_aux.ready) {
        for (int s = 0; s < SLAB_STREAMS; ++s) CK(cudaStreamCreateWithFlags(&c->slab_aux.copy_stream[s], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->slab_aux.fork, cudaEventDisableTiming));
        for (int k = 0; k < SLAB_CHUNKS; ++k)
            for (int s = 0; s < SLAB_STREAMS; ++s) CK(cudaEventCreateWithFlags(&c->slab_aux.chunk[k][s], cudaEventDisableTiming));
        for (int k = 0; k < SLAB_CHUNKS; ++k) CK(cudaEventCreateWithFlags(&c->slab_aux.done[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->slab_aux.pushed, cudaEventDisableTiming));
        c->slab_aux.ready = 1;
    }
    for (int r = 0; r <= ADMP_MAX_PEERS; ++r) {
        const int st = r <= nranks ? (int)((long long)c->K[0] * r / nranks) : c->K[0];
        c->mesh_peers.start[r] = c->spec_peers.start[r] = st;
    }
    c->mesh_peers.n = c->spec_peers.n = nranks;
    c->peer_rank = rank;
    c->peer_n = nranks;
    drop_graph(c);              // a captured SCF body may address the in-place mesh, which a decomposed context does not use
    return 0;
}
static int need_peers(admp_ctx* c) {
    if (need(c, true, false)) return 1;
    if (c->peer_n < 1) return fail("admp_ctx_set_peers has not been called");
    return 0;
}
/* zero the own planes of the mesh */
extern "C" int admp_slab_zero(admp_ctx* c, void* stream) {
    if (need_peers(c)) return 1;
    const size_t plane = (size_t)c->K[1] * c->K[2] * c->w;
    CK(cudaMemsetAsync((char*)c->mesh + plane * c->mesh_peers.start[c->peer_rank], 0, plane * c->mesh_peers.planes(c->peer_rank),
                       (cudaStream_t)stream));
    return 0;
}
/* spread `count` atoms (compact arrays) onto the decomposed mesh */
extern "C" int admp_slab_spread(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride, const void* U,
                                int count) {
    if (need_peers(c)) return 1;
    if (count < 0) return fail("admp_slab_spread: negative count");
    DISPATCH(c, launch_spread, (cudaStream_t)stream, count, c->box, pos, M, M_cols, M_stride, U, c->mesh, &c->mesh_peers);
    CKLAUNCH();
    return 0;
}
/* phase 0: Z-forward + Y-forward (own planes); 1: fused X pass on the own column share (peer access);
 * 2: Y-inverse + Z-inverse (own planes). E_recip / virial sums accumulate this rank's share only. */
extern "C" int admp_slab_fft(admp_ctx* c, void* stream, int phase, int kind, uint32_t flags, double* scalars) {
    if (need_peers(c)) return 1;
    if (phase < 0 || phase > 2) return fail("admp_slab_fft: phase %d", phase);
    fft3d_slab_phase(c->fft, (cudaStream_t)stream, phase, c->peer_rank, c->mesh, c->spec, c->spec_peers, c->box, c->kappa, kind, c->tb,
                     scalars, (flags & ADMP_WANT_VIRIAL) ? 1 : 0, &c->slab_aux);
    CKLAUNCH();
    return 0;
}
/* gather for `count` atoms (compact arrays) from the decomposed potential mesh */
extern "C" int admp_slab_gather(admp_ctx* c, void* stream, const void* pos, const void* M, int M_cols, int M_stride, const void* U,
                                int mode, uint32_t flags, void* dpos, void* G, int G_stride, void* F, double* scalars, int count) {
    if (need_peers(c)) return 1;
    if (count < 0) return fail("admp_slab_gather: negative count");
    DISPATCH(c, launch_gather, (cudaStream_t)stream, count, c->box, pos, M, M_cols, M_stride, U, c->mesh, mode, flags, dpos, G, G_stride, F,
             scalars, &c->mesh_peers);
    CKLAUNCH();
    return 0;
}

/* one Jacobi decision on an already-assembled field F (pair + reciprocal parts, all atoms): adds the
 * self/penalty part, reduces max|F|, tests, updates U (admp/pme.py:133-138). state: int32[8] device,
 * zeroed by the caller before the first cycle; state[5] = continue flag. */
extern "C" int admp_scf_step(admp_ctx* c, void* stream, const void* M, void* U, const void* pol, void* F, int maxiter, double thresh,
                             uint32_t flags, int32_t* state, double* scalars) {
    if (need(c, false, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    cudaGraphConditionalHandle none;
    memset(&none, 0, sizeof(none));
    CK(cudaMemsetAsync(scalars + ADMP_S_MAXFIELD, 0, sizeof(double), st));
    DISPATCH(c, launch_scf_field, st, c->n_atoms, c->kappa, M, U, pol, F, scalars);
    launch_scf_decide(st, state, scalars, maxiter, thresh, (flags & ADMP_WANT_VIRIAL) ? 0 : 1, none, 0);
    DISPATCH(c, launch_scf_update, st, c->n_atoms, state, F, pol, U, 0);
    CKLAUNCH();
    return 0;
}
extern "C" int admp_virial_finalize(admp_ctx* c, void* stream, double* scalars) {
    if (need(c, true, false)) return 1;
    launch_virial_finalize((cudaStream_t)stream, c->box, scalars, c->kvec_ref);
    CKLAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------ fused evaluation
// one pass of optimize_Uind's loop body on staged inputs (admp/pme.py:132-138). The reciprocal chain
// (zero-fill, spread, five FFT passes, field gather) and the real-space pair field are independent: the
// pair kernel runs on a forked stream (a parallel branch of the captured graph) and both accumulate dE/dU
// into Fscf with atomics (Fscf is zero on entry: initial memset, then scf_update_kernel re-zeroes it).
// SCF cycles never accumulate the k-space virial sums (Coulomb QUICK path of the fused X pass); when the
// caller wants the virial, admp_pme_eval runs one more reciprocal pass after the loop, which also serves
// as the refresh pass after the last update (refresh_in_loop = 0).
static int scf_body(admp_ctx* c, cudaStream_t st, int maxiter, double thresh, uint32_t flags, cudaGraphConditionalHandle h, int use_h) {
    const int refresh_in_loop = (flags & ADMP_WANT_VIRIAL) ? 0 : 1;
    CK(cudaEventRecord(c->ev_fork, st));
    CK(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    DISPATCH(c, launch_pme_pair, c->side_stream, c->pairs_cap, c->n_atoms, c->box, c->kappa, c->s_pos, c->s_pairs, c->s_sidx, c->cov_off, c->cov_idx,
             c->cov_nb, c->M, c->s_U, c->s_pol, c->s_th, c->s_mS, c->s_pS, 1, 0u, nullptr, nullptr, c->Fscf, nullptr, nullptr, c->scal, c->rec,
             c->cw.state);
    DISPATCH(c, launch_pme_cluster, c->side_stream, c->cw.n_clusters, c->box, c->kappa, c->cw, c->rec, c->s_U, c->s_mS, c->s_pS, 1, 0u, nullptr,
             nullptr, c->Fscf, nullptr, nullptr, c->scal);
    if (c->phi != nullptr && c->use_custom_fft) {
        // Big meshes (HBM-bound passes): the mesh is dead once the Z-forward pass has read it, so its zero-fill for the NEXT
        // cycle runs on the side stream behind the pair kernel, concurrently with the Y / X / Y passes (which only touch the
        // spectrum; the X pass is FP64-bound and leaves the memory system idle), and the inverse Z pass writes the potential
        // to a second buffer. Entry invariant: c->mesh is zero (admp_pme_eval clears it before the loop).
        DISPATCH(c, launch_spread, st, c->n_atoms, c->box, c->s_pos, c->M, 10, 10, c->s_U, c->mesh);
        fft3d_convolve_roundtrip(c->fft, st, c->mesh, c->spec, c->box, c->kappa, ADMP_CK_COULOMB, c->tb, c->scal, 0, c->phi, c->ev_zfwd);
        CK(cudaStreamWaitEvent(c->side_stream, c->ev_zfwd, 0));
        CK(cudaMemsetAsync(c->mesh, 0, c->mesh_bytes, c->side_stream));
        CK(cudaEventRecord(c->ev_join, c->side_stream));
        DISPATCH(c, launch_gather, st, c->n_atoms, c->box, c->s_pos, c->M, 10, 10, c->s_U, c->phi, 1, 0u, nullptr, nullptr, 10, c->Fscf, c->scal);
    } else {
        CK(cudaEventRecord(c->ev_join, c->side_stream));
        if (recip_field(c, st, c->s_pos, c->M, 10, 10, c->s_U, ADMP_CK_COULOMB, c->scal, 0, false)) return 1;
        DISPATCH(c, launch_gather, st, c->n_atoms, c->box, c->s_pos, c->M, 10, 10, c->s_U, c->phi_cur, 1, 0u, nullptr, nullptr, 10, c->Fscf, c->scal,
                 nullptr, c->phi_zld, c->in_flight > 1);
    }
    CK(cudaStreamWaitEvent(st, c->ev_join, 0));
    DISPATCH(c, launch_scf_field, st, c->n_atoms, c->kappa, c->M, c->s_U, c->s_pol, c->Fscf, c->scal);
    if (flags & ADMP_SCF_CG) {
        // beyond the reference: conjugate gradients on the same body (site.cu scf_cg_kernel). Its last pass is always a field
        // evaluation on the final U, so the mesh / reciprocal energy need no refresh whatever the caller does afterwards.
        DISPATCH(c, launch_scf_cg, st, c->n_atoms, c->state, c->scal, c->Fscf, c->s_pol, c->s_U, c->cg, maxiter, thresh, h, use_h);
    } else {
        launch_scf_decide(st, c->state, c->scal, maxiter, thresh, refresh_in_loop, h, use_h);
        DISPATCH(c, launch_scf_update, st, c->n_atoms, c->state, c->Fscf, c->s_pol, c->s_U, 1);
    }
    CKLAUNCH();
    return 0;
}

// device-resident loop: a CUDA graph whose WHILE node re-runs the body until scf_decide clears it
static int build_scf_graph(admp_ctx* c, int maxiter, double thresh, uint32_t flags) {
    drop_graph(c);
    CK(cudaGraphCreate(&c->graph, 0));
    cudaGraphConditionalHandle handle;
    CK(cudaGraphConditionalHandleCreate(&handle, c->graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = handle;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, c->graph, nullptr, 0, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(c->cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    const int rc = scf_body(c, c->cap_stream, maxiter, thresh, flags, handle, 1);
    cudaGraph_t got = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->cap_stream, &got);
    if (rc) return 1;
    if (e != cudaSuccess) return fail("scf graph capture failed: %s", cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&c->gexec, c->graph, 0));
    c->g_maxiter = maxiter;
    c->g_thresh = thresh;
    c->g_flags = flags;
    return 0;
}

static int run_scf(admp_ctx* c, cudaStream_t st, int maxiter, double thresh, uint32_t flags) {
    CK(cudaMemsetAsync(c->state, 0, sizeof(int32_t) * 8, st));
    const uint32_t gkey = flags & (ADMP_WANT_VIRIAL | ADMP_SCF_CG);
    if ((flags & ADMP_SCF_CG) && !c->cg) {
        drop_graph(c);
        CK(cudaMalloc((void**)&c->cg, sizeof(double) * ((size_t)9 * c->n_atoms + 1)));
    }
    if (!(flags & ADMP_SCF_HOSTSYNC) && !c->graph_failed) {
        if (!c->gexec || c->g_maxiter != maxiter || c->g_thresh != thresh || c->g_flags != gkey) {
            if (build_scf_graph(c, maxiter, thresh, gkey)) {
                c->graph_failed = true;       // keep the message; fall through to the host-synchronised loop
                drop_graph(c);
                cudaGetLastError();
            }
        }
        if (c->gexec) {
            CK(cudaGraphLaunch(c->gexec, st));
            return 0;
        }
    }
    // debug / fallback: same kernels, loop condition read back every iteration
    cudaGraphConditionalHandle none;
    memset(&none, 0, sizeof(none));
    const int max_pass = (flags & ADMP_SCF_CG) ? 2 * maxiter + 3 : maxiter;     // CG: restarts re-evaluate the true residual
    for (int it = 0; it <= max_pass; ++it) {
        if (scf_body(c, st, maxiter, thresh, gkey, none, 0)) return 1;
        CK(cudaMemcpyAsync(c->h_state, c->state, sizeof(int32_t) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!c->h_state[5]) break;
    }
    return 0;
}

static int ensure_pairs(admp_ctx* c, int64_t n_rows) {
    if (n_rows <= c->pairs_cap) return 0;
    drop_graph(c);
    dfree(c->s_pairs);
    dfree(c->s_sidx);
    dfree(c->cw.ent_i);
    dfree(c->cw.ent_m);
    int64_t cap = n_rows + n_rows / 4 + 1024;
    CK(cudaMalloc(&c->s_pairs, sizeof(int32_t) * 2 * cap));
    CK(cudaMalloc(&c->s_sidx, (size_t)cap));
    CK(cudaMalloc(&c->cw.ent_i, sizeof(int32_t) * cap));
    CK(cudaMalloc(&c->cw.ent_m, sizeof(uint32_t) * cap));
    c->pairs_cap = cap;
    return 0;
}

extern "C" int admp_pme_eval(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                             const void* Ql, void* U_io, const void* pol, const void* tholes, const void* mScales,
                             const void* pScales, uint32_t flags, int maxiter, double thresh, double* scalars, void* dpos,
                             void* dQl, void* F, void* dpol, void* dtholes, int32_t* scf_out) {
    if (need(c, true, true)) return 1;
    const bool polz = (pol != nullptr);
    if (polz && (!tholes || !pScales || !U_io)) return fail("admp_pme_eval: polarizable call needs U, pol, tholes, pScales");
    if (polz && c->lmax < 1) return fail("admp_pme_eval: lpol with lmax = 0 is not supported (admp/pme.py:224-228 is broken upstream)");
    if ((flags & ADMP_WANT_GRAD) && (!dpos || !dQl)) return fail("admp_pme_eval: gradient outputs missing");
    if (!scalars) return fail("admp_pme_eval: scalars missing");
    if (polz && (flags & ADMP_SCF) && maxiter < 1) return fail("admp_pme_eval: maxiter must be >= 1 (got %d)", maxiter);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    const int n = c->n_atoms;
    const size_t w = c->w;
    const int nh = (c->lmax + 1) * (c->lmax + 1);
    if (ensure_pairs(c, n_rows)) return 1;
    // stage inputs so that the SCF graph only ever sees context-owned addresses
    CK(cudaMemcpyAsync(c->s_pos, pos, (size_t)n * 3 * w, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(c->s_box, box, 9 * w, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(c->s_pairs, 0, sizeof(int32_t) * 2 * c->pairs_cap, st));        // (0,0) rows are skipped (i<j fails)
    if (n_rows > 0) CK(cudaMemcpyAsync(c->s_pairs, pairs, sizeof(int32_t) * 2 * n_rows, cudaMemcpyDeviceToDevice, st));
    launch_pair_scale(st, c->pairs_cap, n, c->s_pairs, c->cov_off, c->cov_idx, c->cov_nb, c->s_sidx);
    launch_cluster_prepare(st, c->pairs_cap, n, c->cw.n_clusters, c->s_pairs, c->s_sidx, c->cw, c->cluster_force);
    CK(cudaMemcpyAsync(c->s_mS, mScales, 5 * w, cudaMemcpyDeviceToDevice, st));
    if (polz) {
        CK(cudaMemcpyAsync(c->s_U, U_io, (size_t)n * 3 * w, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(c->s_pol, pol, (size_t)n * w, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(c->s_th, tholes, (size_t)n * w, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(c->s_pS, pScales, 5 * w, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaMemsetAsync(c->scal, 0, sizeof(double) * ADMP_S_COUNT, st));
    CK(cudaMemsetAsync(c->G, 0, (size_t)n * 10 * w, st));
    if (dpos) CK(cudaMemsetAsync(dpos, 0, (size_t)n * 3 * w, st));
    if (F) CK(cudaMemsetAsync(F, 0, (size_t)n * 3 * w, st));
    if (dpol) CK(cudaMemsetAsync(dpol, 0, (size_t)n * w, st));
    if (dtholes) CK(cudaMemsetAsync(dtholes, 0, (size_t)n * w, st));
    DISPATCH(c, launch_box_setup, st, c->s_box, c->box, c->K[0], c->K[1], c->K[2]);
    DISPATCH(c, launch_frames_fwd, st, n, c->lmax, c->box, c->s_pos, c->axis_type, c->axis_idx, Ql, c->M, nullptr, nullptr);
    conv_tables(c, st);
    spread_prepare(c, st, c->s_pos);
    CKLAUNCH();
    const int want_vir = (flags & ADMP_WANT_VIRIAL) ? 1 : 0;
    if (polz && (flags & ADMP_SCF)) {
        CK(cudaMemsetAsync(c->Fscf, 0, (size_t)n * 3 * w, st));
        // packed records of the cluster field kernel (it reads the current U from s_U, everything else from here)
        DISPATCH(c, launch_pair_pack, st, n, c->s_pos, c->M, c->s_U, c->s_pol, c->s_th, c->rec);
        const bool two_mesh = c->phi != nullptr && c->use_custom_fft;
        if (two_mesh) CK(cudaMemsetAsync(c->mesh, 0, c->mesh_bytes, st));
        if (run_scf(c, st, maxiter, thresh, flags)) return 1;
        c->phi_cur = two_mesh ? c->phi : mesh_buf(c);
        c->phi_zld = two_mesh ? 0 : mesh_zld(c);
        if (want_vir) {
            // final reciprocal pass on the converged / last-updated U with the k-space virial sums
            CK(cudaMemsetAsync(c->scal + ADMP_S_E_RECIP, 0, sizeof(double), st));
            CK(cudaMemsetAsync(c->scal + ADMP_S_TK, 0, 6 * sizeof(double), st));
            if (recip_field(c, st, c->s_pos, c->M, 10, 10, c->s_U, ADMP_CK_COULOMB, c->scal, 1, false)) return 1;
        }
        if (scf_out) CK(cudaMemcpyAsync(scf_out, c->state + 3, sizeof(int32_t) * 2, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(U_io, c->s_U, (size_t)n * 3 * w, cudaMemcpyDeviceToDevice, st));
    } else {
        if (recip_field(c, st, c->s_pos, c->M, 10, 10, polz ? c->s_U : nullptr, ADMP_CK_COULOMB, c->scal, want_vir, false)) return 1;
    }
    // final evaluation at fixed U (Hellmann-Feynman, pme.py:83-85): phi of the last pass is still in c->mesh
    const void* Uf = polz ? c->s_U : nullptr;
    const uint32_t f = flags & (ADMP_WANT_GRAD | ADMP_WANT_VIRIAL | ADMP_WANT_PGRAD);
    if (flags & ADMP_WANT_GRAD) {
        DISPATCH(c, launch_gather, st, n, c->box, c->s_pos, c->M, 10, 10, Uf, c->phi_cur, 0, f, dpos, c->G, 10, polz ? F : nullptr, c->scal, nullptr,
                 c->phi_zld, c->in_flight > 1);
    }
    DISPATCH(c, launch_pme_pair, st, c->pairs_cap, n, c->box, c->kappa, c->s_pos, c->s_pairs, c->s_sidx, c->cov_off, c->cov_idx, c->cov_nb, c->M, Uf,
             polz ? c->s_pol : nullptr, polz ? c->s_th : nullptr, c->s_mS, polz ? c->s_pS : nullptr, 0, f, dpos, c->G, F, dpol, dtholes, c->scal,
             c->rec, c->cw.state);
    DISPATCH(c, launch_pme_cluster, st, c->cw.n_clusters, c->box, c->kappa, c->cw, c->rec, Uf, c->s_mS, polz ? c->s_pS : nullptr, 0, f, dpos, c->G,
             F, dpol, dtholes, c->scal);
    DISPATCH(c, launch_self, st, n, c->kappa, c->M, Uf, polz ? c->s_pol : nullptr, f, c->G, F, dpol, c->scal);
    if (flags & ADMP_WANT_GRAD) {
        DISPATCH(c, launch_frames_bwd, st, n, c->lmax, c->box, c->s_pos, c->axis_type, c->axis_idx, Ql, c->G, dQl, dpos, c->scal, want_vir);
    }
    if (want_vir) launch_virial_finalize(st, c->box, c->scal, c->kvec_ref);
    CKLAUNCH();
    CK(cudaMemcpyAsync(scalars, c->scal, sizeof(double) * ADMP_S_COUNT, cudaMemcpyDeviceToDevice, st));
    (void)nh;
    return 0;
}

extern "C" int admp_disp_eval(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                              const void* c_list, const void* mScales, int pmax, uint32_t flags, double* scalars, void* dpos,
                              void* dc) {
    if (need(c, true, true)) return 1;
    if (pmax != 6 && pmax != 8 && pmax != 10) return fail("admp_disp_eval: pmax must be 6, 8 or 10");
    if ((flags & ADMP_WANT_GRAD) && !dpos) return fail("admp_disp_eval: dpos missing");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    const int n = c->n_atoms;
    const size_t w = c->w;
    CK(cudaMemsetAsync(c->scal, 0, sizeof(double) * ADMP_S_COUNT, st));
    if (dpos) CK(cudaMemsetAsync(dpos, 0, (size_t)n * 3 * w, st));
    if (dc) CK(cudaMemsetAsync(dc, 0, (size_t)n * 3 * w, st));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0], c->K[1], c->K[2]);
    const uint32_t f = flags & (ADMP_WANT_GRAD | ADMP_WANT_VIRIAL | ADMP_WANT_PGRAD);
    DISPATCH(c, launch_disp_pair, st, n_rows, n, c->box, c->kappa, pmax, pos, pairs, c->cov_off, c->cov_idx, c->cov_nb, c_list, mScales, f,
             dpos, (f & ADMP_WANT_PGRAD) ? dc : nullptr, c->scal);
    CKLAUNCH();
    const int kinds[3] = {ADMP_CK_DISP6, ADMP_CK_DISP8, ADMP_CK_DISP10};
    for (int p = 0; p < (pmax - 4) / 2; ++p) {
        const char* col = (const char*)c_list + p * w;
        if (recip_field(c, st, pos, col, 1, 3, nullptr, kinds[p], c->scal, (f & ADMP_WANT_VIRIAL) ? 1 : 0)) return 1;
        if (f & (ADMP_WANT_GRAD | ADMP_WANT_PGRAD)) {
            void* g = ((f & ADMP_WANT_PGRAD) && dc) ? (void*)((char*)dc + p * w) : nullptr;
            DISPATCH(c, launch_gather, st, n, c->box, pos, col, 1, 3, nullptr, c->phi_cur, 0, f, (f & ADMP_WANT_GRAD) ? dpos : nullptr, g, 3,
                     nullptr, c->scal, nullptr, c->phi_zld, c->in_flight > 1);
            CKLAUNCH();
        }
    }
    DISPATCH(c, launch_disp_self, st, n, c->kappa, pmax, c_list, f, dc, c->scal);
    if (f & ADMP_WANT_VIRIAL) launch_virial_finalize(st, c->box, c->scal, c->kvec_ref);
    CKLAUNCH();
    CK(cudaMemcpyAsync(scalars, c->scal, sizeof(double) * ADMP_S_COUNT, cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int tt_pair_impl(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                        const void* mScales, const void* a, const void* b, const void* q, const void* c6, const void* c8, const void* c10,
                        int n_par, uint32_t flags, double* scalars, void* dpos, void* dparams) {
    if (need(c, false, true)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    const int n = c->n_atoms;
    const size_t w = c->w;
    CK(cudaMemsetAsync(c->scal, 0, sizeof(double) * ADMP_S_COUNT, st));
    if (dpos) CK(cudaMemsetAsync(dpos, 0, (size_t)n * 3 * w, st));
    if (dparams) CK(cudaMemsetAsync(dparams, 0, (size_t)n * n_par * w, st));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0] ? c->K[0] : 6, c->K[1] ? c->K[1] : 6, c->K[2] ? c->K[2] : 6);
    const uint32_t f = flags & (ADMP_WANT_GRAD | ADMP_WANT_VIRIAL | ADMP_WANT_PGRAD);
    DISPATCH(c, launch_tt_pair, st, n_rows, n, c->box, pos, pairs, c->cov_off, c->cov_idx, c->cov_nb, mScales, a, b, q, c6, c8, c10, f, dpos,
             (f & ADMP_WANT_PGRAD) ? dparams : nullptr, c->scal);
    CKLAUNCH();
    CK(cudaMemcpyAsync(scalars, c->scal, sizeof(double) * ADMP_S_COUNT, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int admp_tt_pair(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                            const void* mScales, const void* a, const void* b, const void* q, const void* cc, uint32_t flags,
                            double* scalars, void* dpos, void* dparams) {
    return tt_pair_impl(c, stream, pos, box, pairs, n_rows, mScales, a, b, q, cc, nullptr, nullptr, 4, flags, scalars, dpos, dparams);
}

extern "C" int admp_tt_pair_c10(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                                const void* mScales, const void* a, const void* b, const void* q, const void* c6, const void* c8,
                                const void* c10, uint32_t flags, double* scalars, void* dpos, void* dparams) {
    if (!c6 || !c8 || !c10) return fail("admp_tt_pair_c10: c6, c8 and c10 are required");
    return tt_pair_impl(c, stream, pos, box, pairs, n_rows, mScales, a, b, q, c6, c8, c10, 6, flags, scalars, dpos, dparams);
}

static int nblist_build_impl(admp_ctx* c, cudaStream_t st, const void* pos, const void* box, const double* hb, int n, double rc,
                             int32_t* pairs, int64_t capacity, int32_t* info);

extern "C" int admp_pair_geometry(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                                  void* dr, int32_t* sidx) {
    if (need(c, false, true)) return 1;
    if (!dr || !sidx) return fail("admp_pair_geometry: outputs missing");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0] ? c->K[0] : 6, c->K[1] ? c->K[1] : 6, c->K[2] ? c->K[2] : 6);
    DISPATCH(c, launch_pair_geom, st, n_rows, c->n_atoms, c->box, pos, pairs, c->cov_off, c->cov_idx, c->cov_nb, dr, sidx);
    CKLAUNCH();
    return 0;
}

extern "C" int admp_pair_geometry_bwd(admp_ctx* c, void* stream, const void* pos, const void* box, const int32_t* pairs, int64_t n_rows,
                                      const void* g_dr, uint32_t flags, void* dpos, double* scalars) {
    if (need(c, false, true)) return 1;
    if (!g_dr || !dpos || !scalars) return fail("admp_pair_geometry_bwd: arguments missing");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(scalars, 0, sizeof(double) * ADMP_S_COUNT, st));
    CK(cudaMemsetAsync(dpos, 0, (size_t)c->n_atoms * 3 * c->w, st));
    DISPATCH(c, launch_box_setup, st, box, c->box, c->K[0] ? c->K[0] : 6, c->K[1] ? c->K[1] : 6, c->K[2] ? c->K[2] : 6);
    DISPATCH(c, launch_pair_geom_bwd, st, n_rows, c->n_atoms, c->box, pos, pairs, g_dr, flags, dpos, scalars);
    CKLAUNCH();
    return 0;
}

extern "C" int admp_nblist_build(admp_ctx* c, void* stream, const void* pos, const void* box, int n, double rc, int32_t* pairs,
                                 int64_t capacity, int32_t* info) {
    if (!c) return fail("null ctx");
    if (n <= 0 || rc <= 0.0) return fail("admp_nblist_build: bad n_atoms / rc");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(c->device));
    // cell grid from the box lengths: needs the box on the host once (the build is the only
    // stage whose launch geometry depends on the box)
    double hb[9];
    if (c->dtype == ADMP_F64) CK(cudaMemcpyAsync(hb, box, 72, cudaMemcpyDeviceToHost, st));
    else {
        float fb[9];
        CK(cudaMemcpyAsync(fb, box, 36, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < 9; ++k) hb[k] = fb[k];
    }
    CK(cudaStreamSynchronize(st));
    return nblist_build_impl(c, st, pos, box, hb, n, rc, pairs, capacity, info);
}

// same, with the caller's host copy of the box (9 doubles, row-major): no device-to-host read, no host synchronisation
extern "C" int admp_nblist_build_hostbox(admp_ctx* c, void* stream, const void* pos, const void* box, const double* box_host, int n,
                                         double rc, int32_t* pairs, int64_t capacity, int32_t* info) {
    if (!c) return fail("null ctx");
    if (!box_host) return fail("admp_nblist_build_hostbox: null host box");
    if (n <= 0 || rc <= 0.0) return fail("admp_nblist_build: bad n_atoms / rc");
    CK(cudaSetDevice(c->device));
    return nblist_build_impl(c, (cudaStream_t)stream, pos, box, box_host, n, rc, pairs, capacity, info);
}

static int nblist_build_impl(admp_ctx* c, cudaStream_t st, const void* pos, const void* box, const double* hb, int n, double rc,
                             int32_t* pairs, int64_t capacity, int32_t* info) {
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            if (a != b && hb[3 * a + b] != 0.0) return fail("admp_nblist_build: orthorhombic boxes only");
    int nc[3];
    long long ncell = 1;
    for (int d = 0; d < 3; ++d) {
        if (!(hb[4 * d] > 0.0)) return fail("admp_nblist_build: non-positive box length");
        if (2.0 * rc > hb[4 * d]) return fail("admp_nblist_build: rc exceeds half the box length (minimum image)");
        int v = (int)std::floor(hb[4 * d] / (rc * (1.0 + 1e-6)));
        if (v < 1) v = 1;
        if (v > 512) v = 512;
        nc[d] = v;
        ncell *= v;
    }
    if (n > c->nb.capacity_atoms) {
        dfree(c->nb.cell_of); dfree(c->nb.sorted); dfree(c->nb.nbr_count); dfree(c->nb.nbr_start);
        CK(cudaMalloc(&c->nb.cell_of, sizeof(int32_t) * n));
        CK(cudaMalloc(&c->nb.sorted, sizeof(int32_t) * n));
        CK(cudaMalloc(&c->nb.nbr_count, sizeof(int32_t) * (n + 1)));
        CK(cudaMalloc(&c->nb.nbr_start, sizeof(int32_t) * (n + 1)));
        c->nb.capacity_atoms = n;
    }
    if (ncell > c->nb.capacity_cells) {
        dfree(c->nb.cell_count); dfree(c->nb.cell_start);
        CK(cudaMalloc(&c->nb.cell_count, sizeof(int32_t) * (ncell + 1)));
        CK(cudaMalloc(&c->nb.cell_start, sizeof(int32_t) * (ncell + 1)));
        c->nb.capacity_cells = (int)ncell;
    }
    const int K1 = c->K[0] ? c->K[0] : 6, K2 = c->K[1] ? c->K[1] : 6, K3 = c->K[2] ? c->K[2] : 6;
    DISPATCH(c, launch_box_setup, st, box, c->box, K1, K2, K3);
    CK(launch_nblist(st, c->box, pos, c->dtype, n, rc, c->nb, nc[0], nc[1], nc[2], pairs, capacity, info));
    return 0;
}
