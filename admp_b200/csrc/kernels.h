// Internal launcher declarations shared by the translation units of libadmp_b200.
#pragma once
#include "common.cuh"
#include "influence.cuh"

namespace admp {

// frames.cu
template <typename T>
void launch_frames_fwd(cudaStream_t st, int n, int lmax, const BoxInfo* B, const void* pos, const int32_t* atype,
                       const int32_t* ai, const void* Ql, void* M, void* Qg, void* Fr);
template <typename T>
void launch_frames_bwd(cudaStream_t st, int n, int lmax, const BoxInfo* B, const void* pos, const int32_t* atype,
                       const int32_t* ai, const void* Ql, const void* G, void* dQl, void* dpos, double* scalars, int want_box,
                       int first = 0);

template <typename T> void launch_rotate(cudaStream_t st, int64_t n, int lmax, int to_local, const void* Q, const void* Fr, void* out);

// pair.cu
void launch_pair_scale(cudaStream_t st, int64_t n_rows, int n_atoms, const int32_t* pairs, const int32_t* cov_off,
                       const int32_t* cov_idx, const int8_t* cov_nb, int8_t* sidx);
template <typename T>
void launch_pme_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, double kappa, const void* pos,
                     const int32_t* pairs, const int8_t* sidx, const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb,
                     const void* M, const void* U, const void* pol, const void* tholes, const void* mS, const void* pS,
                     int mode, uint32_t flags, void* dpos, void* G, void* F, void* dpol, void* dth, double* scalars, void* rec,
                     const int32_t* sel = nullptr);      // sel: cluster-kernel selector (ClusterWork::state); the kernel is a no-op when sel[2] != 0
size_t pair_record_bytes(int dtype_bytes);        // per-atom record of the staged pair kernel (workspace `rec`)
template <typename T>
void launch_pair_pack(cudaStream_t st, int n, const void* pos, const void* M, const void* U, const void* pol, const void* tholes, void* rec);

// pair_generic.cu - geometry + adjoint halves of the generic pair driver (any user kernel in between)
template <typename T>
void launch_pair_geom(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                      const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb, void* dr, int32_t* sidx);
template <typename T>
void launch_pair_geom_bwd(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                          const void* gdr, uint32_t flags, void* dpos, double* scalars);

// pair_cluster.cu - j-cluster x i-lane tiles built from the caller's pair rows (dense, (j, i)-sorted lists)
struct ClusterWork {
    int32_t *cl_of, *cl_first, *cl_size;   // per atom / per cluster (host-built from the covalent map at set_topology)
    int32_t* row_start;                    // n_atoms + 1: first row of every j
    int32_t* ent_i;                        // pairs_cap: i atom of every tile entry
    uint32_t* ent_m;                       // pairs_cap: 4 bits per cluster slot: listed, scale index
    int32_t* cl_extra;                     // n_clusters: entries beyond slot 0's own list
    int32_t* state;                        // int32[8]: [1] live rows, [2] cluster kernel selected, [3] group column (pair_cluster.cu)
    int n_clusters, min_rows_per_cluster;
};
void launch_cluster_prepare(cudaStream_t st, int64_t n_rows, int n_atoms, int n_clusters, const int32_t* pairs, const int8_t* sidx,
                            const ClusterWork& w, int force);
template <typename T>
void launch_pme_cluster(cudaStream_t st, int n_clusters, const BoxInfo* B, double kappa, const ClusterWork& w, const void* rec,
                        const void* U, const void* mS, const void* pS, int mode, uint32_t flags, void* dpos, void* G, void* F,
                        void* dpol, void* dth, double* scalars);
template <typename T>
void launch_disp_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, double kappa, int pmax, const void* pos,
                      const int32_t* pairs, const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb,
                      const void* c_list, const void* mS, uint32_t flags, void* dpos, void* dc, double* scalars);
template <typename T>
void launch_tt_pair(cudaStream_t st, int64_t n_rows, int n_atoms, const BoxInfo* B, const void* pos, const int32_t* pairs,
                    const int32_t* cov_off, const int32_t* cov_idx, const int8_t* cov_nb, const void* mS, const void* a,
                    const void* b, const void* q, const void* c, const void* c8, const void* c10, uint32_t flags, void* dpos,
                    void* dparams, double* scalars);

// recip.cu
template <typename T>
void launch_spread(cudaStream_t st, int n, const BoxInfo* B, const void* pos, const void* M, int m_cols, int m_stride,
                   const void* U, void* mesh, const PeerTab* peers = nullptr, int zld = 0);   // zld: reals per mesh line (0: K3)
template <typename T>
void launch_convolve(cudaStream_t st, const BoxInfo* B, size_t n_half, int n_sm, double kappa, int kind, const ConvTables& tb,
                     void* S, double* scalars, int want_vir);
void launch_conv_tables(cudaStream_t st, const BoxInfo* B, double kappa, const double* bt1, const double* bt2, const double* bt3,
                        double* ek, double* k2, int* ortho, int maxK);
template <typename T>
void launch_gather(cudaStream_t st, int n, const BoxInfo* B, const void* pos, const void* M, int m_cols, int m_stride, const void* U,
                   const void* phi, int mode, uint32_t flags, void* dpos, void* G, int g_stride, void* F, double* scalars,
                   const PeerTab* peers = nullptr, int zld = 0, int narrow = 0);

// spread_brick.cu - brick-staged spread (mesh written once, zero-fill included); atoms binned by home brick per evaluation
struct BrickGeom { int nb[3]; int bz; };       // bricks per dimension (16 x 16 x bz points each)
struct BrickWork {
    int32_t* count = nullptr;      // n_bricks + 1: histogram, then fill cursors
    int32_t* start = nullptr;      // n_bricks + 1: first sorted slot of every brick
    int4* tmp = nullptr;           // n_atoms: (i0, j0, k0, home brick) in atom order
    int4* anchor = nullptr;        // n_atoms: (i0, j0, k0, atom) sorted by home brick
    void* rec = nullptr;           // n_atoms x 72 reals: per-atom spline / coefficient records of the current spread (sorted order)
    size_t rec_bytes = 0;
    BrickGeom geom = {};
    int n_bricks = 0, n_atoms = 0;
    bool ready = false;
};
bool brick_supported(const int K[3]);
cudaError_t brick_alloc(BrickWork& w, int n_atoms, const int K[3], int n_sm, size_t elem_bytes);
void brick_free(BrickWork& w);
size_t brick_bytes(const BrickWork& w);
template <typename T> void launch_brick_sort(cudaStream_t st, const BrickWork& w, const BoxInfo* B, const void* pos);
template <typename T>
void launch_spread_brick(cudaStream_t st, const BrickWork& w, const BoxInfo* B, const void* pos, const void* M, int m_cols,
                         int m_stride, const void* U, void* mesh);

// fft.cu - hand-written 3-D real FFT fused with the influence-function convolution
struct Fft3d;
Fft3d* fft3d_create(int K1, int K2, int K3, int dtype, const char** why);     // nullptr when the sizes are unsupported
void fft3d_destroy(Fft3d* f);
void fft3d_forward(Fft3d* f, cudaStream_t st, const void* mesh, void* spec);
void fft3d_inverse(Fft3d* f, cudaStream_t st, void* spec, void* mesh);
void fft3d_single_pass(Fft3d* f, cudaStream_t st, int which, void* mesh, void* spec, const BoxInfo* B, double kappa, int kind,
                       const ConvTables& tb, double* scalars);
void fft3d_convolve_roundtrip(Fft3d* f, cudaStream_t st, void* mesh, void* spec, const BoxInfo* B, double kappa, int kind,
                              const ConvTables& tb, double* scalars, int want_vir, void* mesh_out = nullptr,
                              cudaEvent_t after_zfwd = nullptr, int zld = 0);
// mesh == spec (in place, line by line; zld = 2 (K3/2 + 1) reals per mesh line) is allowed when this returns true
bool fft3d_inplace_supported(const Fft3d* f);

bool fft3d_slab_supported(const Fft3d* f);
constexpr int SLAB_CHUNKS = 8;        // maximum pipeline depth of the pulled X pass (ADMP_SLAB_CHUNKS, default 4)
constexpr int SLAB_STREAMS = 4;       // copy streams the peer pulls of one chunk are spread over (several DMA engines)
struct SlabAux {                      // copy streams + events of the pull pipeline (owned by the context)
    cudaStream_t copy_stream[SLAB_STREAMS];
    cudaEvent_t fork;
    cudaEvent_t chunk[SLAB_CHUNKS][SLAB_STREAMS];
    cudaEvent_t done[SLAB_CHUNKS];      // transform of chunk k finished (its push may start)
    cudaEvent_t pushed;                 // all pushes of the pass issued behind this event
    int ready;
};
void fft3d_slab_phase(Fft3d* f, cudaStream_t st, int phase, int rank, void* mesh, void* spec, const PeerTab& spec_peers,
                      const BoxInfo* B, double kappa, int kind, const ConvTables& tb, double* scalars, int want_vir,
                      const SlabAux* aux);

// site.cu
template <typename T> void launch_box_setup(cudaStream_t st, const void* box, BoxInfo* B, int K1, int K2, int K3);
template <typename T>
void launch_self(cudaStream_t st, int n, double kappa, const void* M, const void* U, const void* pol, uint32_t flags, void* G,
                 void* F, void* dpol, double* scalars);
template <typename T>
void launch_disp_self(cudaStream_t st, int n, double kappa, int pmax, const void* c_list, uint32_t flags, void* dc, double* scalars);
template <typename T>
void launch_scf_field(cudaStream_t st, int n, double kappa, const void* M, const void* U, const void* pol, void* F, double* scalars);
void launch_scf_decide(cudaStream_t st, int32_t* state, double* scalars, int maxiter, double thresh, int refresh_in_loop,
                       cudaGraphConditionalHandle handle, int use_handle);
template <typename T>
void launch_scf_update(cudaStream_t st, int n, const int32_t* state, void* F, const void* pol, void* U, int zero_F);
// preconditioned conjugate gradients on the Jacobi fixed point (beyond the reference): replaces decide + update
template <typename T>
void launch_scf_cg(cudaStream_t st, int n, int32_t* state, double* scalars, void* F, const void* pol, void* Us, double* cg, int maxiter,
                   double thresh, cudaGraphConditionalHandle handle, int use_handle);
void launch_virial_finalize(cudaStream_t st, const BoxInfo* B, double* scalars, int kvec_ref);

// nblist.cu
struct NbWork {
    int32_t* cell_of;      // n
    int32_t* cell_count;   // ncell+1
    int32_t* cell_start;   // ncell+1
    int32_t* sorted;       // n
    int32_t* nbr_count;    // n+1
    int32_t* nbr_start;    // n+1
    int capacity_atoms, capacity_cells;
    void* geom;            // NbGeom (device), per context
    void* scan_tmp;        // CUB scan scratch, per context
    size_t scan_tmp_bytes;
};
cudaError_t launch_nblist(cudaStream_t st, const BoxInfo* B, const void* pos, int dtype, int n, double rc, NbWork& w, int ncx, int ncy,
                   int ncz, int32_t* pairs, int64_t capacity, int32_t* info);

}  // namespace admp
