/*
 * admp_b200.h - C ABI of libadmp_b200.so: B200 (sm_100a) kernels for ADMP's multipolar
 * PME hot path.  Plain pointers and sizes only; no torch / jax types.
 *
 * The reference (Roy-Kid/ADMP) is pure Python on JAX and has no FFI of its own; the
 * entry points below are what a jax.ffi / ctypes binding for this path binds to, one
 * per reference function (cited as admp/<file>:<line>).  INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; admp_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - `stream` is a cudaStream_t passed as void*; compute calls never synchronise the
 *     host, never allocate, and are CUDA-graph capturable (workspaces live in the ctx
 *     and are sized by admp_ctx_set_* calls);
 *   - all array arguments of compute calls are DEVICE pointers; `real` means double
 *     when the ctx dtype is ADMP_F64 and float when ADMP_F32 (settings.PRECISION,
 *     admp/settings.py:5); scalar accumulators (energies, virial, scale gradients) are
 *     always double;
 *   - units: Angstrom, e, kJ/mol (DIELECTRIC = 1389.35455846, admp/pme.py:16);
 *   - box is 3x3 row-major, rows = lattice vectors (admp/spatial.py:13-32);
 *   - harmonic multipole order 00,10,11c,11s,20,21c,21s,22c,22s (admp/multipole.py);
 *   - internal Cartesian site layout ("M", 10 reals per site):
 *         q, mu_x, mu_y, mu_z, T_xx, T_xy, T_xz, T_yy, T_yz, T_zz
 *     with T the traceless Cartesian quadrupole of Stone's convention; gradients w.r.t.
 *     M ("G") use the same layout, one variable per off-diagonal element.
 */
#ifndef ADMP_B200_H
#define ADMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADMP_F64 0
#define ADMP_F32 1

/* reciprocal-space influence functions, admp/recip.py:434-462 */
#define ADMP_CK_COULOMB 1
#define ADMP_CK_DISP6 6
#define ADMP_CK_DISP8 8
#define ADMP_CK_DISP10 10

/* slots of the double-precision scalar block ("scalars") every compute call accumulates into */
#define ADMP_S_E_REAL 0    /* pme_real / disp_pme_real / pair_int energy */
#define ADMP_S_E_RECIP 1   /* pme_recip energy (last convolution) */
#define ADMP_S_E_SELF 2    /* pme_self / disp_pme_self */
#define ADMP_S_E_PEN 3     /* pol_penalty */
#define ADMP_S_DBOX 4      /* 9 slots: dE/dbox[a][b] from real space + frames (image shifts) */
#define ADMP_S_DNSTAR 13   /* 9 slots: dE/dNstar[d][c] from spread/gather */
#define ADMP_S_TK 22       /* 6 slots: sum_k dC/dk2 |S|^2/theta^2 k_a k_c (xx,xy,xz,yy,yz,zz) */
#define ADMP_S_DMSCALE 28  /* 5 slots: dE/dmScales */
#define ADMP_S_DPSCALE 33  /* 5 slots: dE/dpScales */
#define ADMP_S_MAXFIELD 38 /* max |dE/dU| over polarizable sites (bit pattern of a double) */
#define ADMP_S_COUNT 48

/* flags for admp_pme_eval */
#define ADMP_WANT_GRAD 1u     /* dE/dpositions, dE/dQ_local, dE/dU */
#define ADMP_WANT_VIRIAL 2u   /* dE/dbox */
#define ADMP_WANT_PGRAD 4u    /* dE/dmScales, dE/dpScales, dE/dtholes, dE/dpol */
#define ADMP_SCF 8u           /* run optimize_Uind before the final evaluation */
#define ADMP_SCF_HOSTSYNC 16u /* debug: host-synchronised SCF loop instead of the device-resident graph */
#define ADMP_SCF_CG 32u       /* beyond the reference: converge U with conjugate gradients preconditioned by pol/DIELECTRIC instead of
                               * the Jacobi iteration of admp/pme.py:132-138 (same fixed point and stopping rule, evaluated on the
                               * final U; scf_out[0] = CG iterations). Default off: the reference's iteration, cycle for cycle. */
/* admp_pme_real only: the pair rows, topology and scales index are those of the previous admp_pme_real call on this
 * context - reuse its per-row scale indices and cluster tiles (what admp_pme_eval does for the 31 pair passes of one
 * polarizable evaluation) instead of rebuilding them */
#define ADMP_REUSE_PAIR_TILES 0x20000000u

typedef struct admp_ctx admp_ctx;

const char* admp_last_error(void);
int admp_version(void);

/* ---- context / environment: ADMPPmeForce.__init__, update_env (admp/pme.py:37-55, :89-94) */
int admp_ctx_create(admp_ctx** out, int device, int dtype);
int admp_ctx_destroy(admp_ctx* ctx);
/* kappa, K1..K3 are independent env values. lmax in {0,1,2}. Re-plans cuFFT. */
int admp_ctx_set_pme(admp_ctx* ctx, double kappa, int K1, int K2, int K3, int lmax);
/* HOST pointers, copied: axis types / anchor indices (admp/spatial.py:44-74) and the sparse
 * covalent map (CSR over atoms; replaces the dense Na x Na matrix of admp/pme.py:681). Either
 * group may be NULL (no multipole frames / no bonded pairs). */
int admp_ctx_set_topology(admp_ctx* ctx, int n_atoms, const int32_t* axis_type,
                          const int32_t* axis_indices, const int32_t* cov_offsets,
                          const int32_t* cov_index, const int8_t* cov_nbonds);
/* k-vector convention of the k-space virial term. 0 (default): component i of k belongs to mesh axis i (the
 * chain-rule-correct dE/dbox). 1: the reference's table, built with meshgrid(kz, kx, ky) (admp/recip.py:339-341),
 * which exchanges the roles of axes 0 and 1 in dk^2/dbox; on a cubic cell with K1 = K2 = K3 (the only case in which
 * the reference is self-consistent) this reproduces the diagonal of jax.grad(energy, argnums=box) entry by entry.
 * Energies, forces and every other gradient do not depend on it. */
int admp_ctx_set_kvec_order(admp_ctx* ctx, int reference);
/* Real-space pair traversal of admp_pme_eval. Two kernels evaluate the caller's pair rows (identical pair set, identical
 * arithmetic): the flat one (one row per thread, any row order) and the cluster one (a warp per j-cluster = bonded group of
 * <= 4 consecutive atoms, lanes = the union of the cluster's i neighbours; needs rows grouped by j with ascending i - the order
 * admp_nblist_build / jax_md's OrderedSparse produce). force: 0 = choose on the device (cluster when the rows are ordered and the
 * list holds >= min_rows_per_cluster rows per cluster, default 96), 1 = cluster whenever the order allows, -1 = flat.
 * admp_ctx_pair_cluster_active reads back which one the last evaluation used (1 cluster, 0 flat; synchronises - tests only). */
int admp_ctx_set_pair_cluster(admp_ctx* ctx, int force, int min_rows_per_cluster);
/* Hint: the caller keeps n evaluations in flight (n contexts on n streams - the frame batches of force-field fitting,
 * admp/api.py usage in examples/openmm_api/run.py:41-45). n > 1 selects kernel shapes that leave SM room for the other
 * evaluations (today: the narrow gather); results are unchanged up to summation order. Default 1. */
int admp_ctx_set_in_flight(admp_ctx* ctx, int n);
int admp_ctx_pair_cluster_active(admp_ctx* ctx);
/* B-spline spread of admp_pme_eval / admp_pme_recip / admp_pme_spread (replaces admp/recip.py:313-392). Two kernels, same
 * result to rounding. bricks = 0 (default): zero-fill + one warp per atom with global atomics. bricks = 1: brick-staged - atoms
 * binned by 16 x 16 x 16|32 mesh bricks once per evaluation, one block accumulates a brick in shared memory (no atomics) and
 * writes it once with coalesced stores (the zero-fill is part of the write); needs >= 3 bricks per mesh dimension, not
 * available for x-slab decomposed meshes / atom sub-ranges. Measured slower on B200 on every workload (each atom is visited
 * by ~2 bricks and the kernel is instruction-bound; DESIGN.md section 5, profiles/r2i_brick_spread.md), hence opt-in.
 * admp_ctx_spread_bricks returns the brick depth in z (16 / 32) when the brick path is active, 0 otherwise.
 * Environment: ADMP_SPREAD=bricks selects it at context creation, ADMP_BRICK_Z=16|32 forces the depth. */
int admp_ctx_set_spread(admp_ctx* ctx, int bricks);
int admp_ctx_spread_bricks(const admp_ctx* ctx);
int64_t admp_ctx_workspace_bytes(const admp_ctx* ctx);
/* 1 when optimize_Uind runs as the device-resident CUDA-graph WHILE loop, 0 when the
 * host-synchronised loop is in use (ADMP_SCF_HOSTSYNC or graph construction failed). */
int admp_ctx_scf_graph_active(const admp_ctx* ctx);
/* FFT backend: 1 = hand-written mixed-radix FFT fused with the convolution (mesh sizes with prime
 * factors in {2,3,5,7,11,13}, K3 even), 0 = cuFFT + separate convolution kernel. ADMP_FFT=cufft in
 * the environment selects cuFFT at set_pme time. */
int admp_ctx_fft_backend(const admp_ctx* ctx);
int admp_ctx_set_fft_backend(admp_ctx* ctx, int custom);

/* ---- stage entry points (each mirrors one reference function) ------------------------- */

/* generate_construct_local_frames + rot_local2global (admp/spatial.py:76-142,
 * admp/multipole.py:183-201). Outputs (any may be NULL): M (n,10), Qg (n,9) harmonic,
 * frames (n,9). */
int admp_frames_fwd(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                    const void* Q_local, void* M, void* Qg, void* frames);
/* adjoint of the above: G (n,10) -> dQ_local (n,9) written, dpos (n,3) and
 * scalars[ADMP_S_DBOX..] accumulated. */
int admp_frames_bwd(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                    const void* Q_local, const void* G, void* dQ_local, void* dpos,
                    double* scalars);

/* rot_global2local (to_local = 1) / rot_local2global (to_local = 0) with explicit frames
 * (admp/multipole.py:92-201): Q, out (n,(lmax+1)^2) harmonic; frames (n,9) rows = axes. */
int admp_rotate(admp_ctx* ctx, void* stream, int64_t n, int lmax, int to_local, const void* Q,
                const void* frames, void* out);

/* pme_real + pme_real_kernel + calc_e_perm + calc_e_ind (admp/pme.py:258-729), fused with
 * every adjoint. pairs: (n_rows,2) int32, rows with p0<p1 are evaluated (pme.py:671).
 * U/pol/tholes/pScales NULL => non-polarizable. mode 0: energy + adjoints, 1: dE/dU only. */
int admp_pme_real(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                  const int32_t* pairs, int64_t n_rows, const void* M, const void* U,
                  const void* pol, const void* tholes, const void* mScales, const void* pScales,
                  int mode, uint32_t flags, void* dpos, void* G, void* F, void* dpol,
                  void* dtholes, double* scalars);

/* generate_pme_recip / pme_recip (admp/recip.py:21-431): spread -> cuFFT -> influence
 * function (kind = ADMP_CK_*) + energy -> cuFFT -> gather. M_cols = 10 (multipoles) or 1
 * (lmax = 0: charges / dispersion coefficients, stride `M_stride` reals). U may be NULL. */
int admp_pme_recip(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                   const void* M, int M_cols, int M_stride, const void* U, int kind, int mode,
                   uint32_t flags, void* dpos, void* G, int G_stride, void* F, double* scalars);

/* the individual stages admp_pme_recip chains, operating on the context's mesh / spectrum
 * (recip.py:368-392 spread_Q; :410 fftn; :400-426 influence function + energy; gather = adjoint of
 * spread; mode 1 = field only, ACCUMULATED into F with atomics). admp_pme_spread zero-fills then scatters;
 * _spread_only is the bare scatter kernel. */
int admp_pme_spread(admp_ctx* ctx, void* stream, const void* pos, const void* box, const void* M,
                    int M_cols, int M_stride, const void* U);
int admp_pme_spread_only(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols,
                         int M_stride, const void* U);
int admp_pme_fft(admp_ctx* ctx, void* stream, int inverse);
int admp_pme_convolve(admp_ctx* ctx, void* stream, int kind, uint32_t flags, double* scalars);
/* fused five-pass round trip mesh -> phi (Z-fwd, Y-fwd, [X-fwd * C_k/theta^2 * X-inv], Y-inv, Z-inv)
 * of the hand-written FFT; admp_pme_recip / admp_pme_eval use it whenever the mesh sizes allow. */
int admp_pme_fft_convolve(admp_ctx* ctx, void* stream, int kind, uint32_t flags, double* scalars);
/* one pass of that round trip: 0 Z-fwd, 1 Y-fwd, 2 fused X, 3 Y-inv, 4 Z-inv (profiling / roofline timing) */
int admp_pme_fft_pass(admp_ctx* ctx, void* stream, int which, int kind, double* scalars);
int admp_pme_gather(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols,
                    int M_stride, const void* U, int mode, uint32_t flags, void* dpos, void* G,
                    int G_stride, void* F, double* scalars);
/* which: 0 = real mesh (K1*K2*K3 reals), 1 = half spectrum (K1*K2*(K3/2+1) complex)
 * Buffer 0 is the mesh of the stage entry points above and of the decompositions. The fused evaluations (admp_pme_eval,
 * admp_disp_eval, admp_pme_recip) keep their real mesh INSIDE buffer 1 (line l of K3 reals at the start of spectrum line l, the layout
 * of an in-place real-to-complex transform) whenever the tile-pipelined Z passes serve the mesh size; ADMP_MESH_INPLACE=0 makes them
 * use buffer 0. Neither buffer holds anything a caller may rely on after a fused evaluation. */
void* admp_ctx_buffer(admp_ctx* ctx, int which);
/* device-to-device copy between a caller buffer and the context's mesh / spectrum (tests, tools) */
int admp_ctx_buffer_io(admp_ctx* ctx, void* stream, int which, void* user, int64_t nbytes, int to_ctx);

/* pme_self + pol_penalty (admp/pme.py:738-774). */
int admp_pme_self(admp_ctx* ctx, void* stream, const void* M, const void* U, const void* pol,
                  uint32_t flags, void* G, void* F, void* dpol, double* scalars);

/* ---- atom-range stages: building blocks of the multi-GPU atom-block decomposition -------------
 * (admp_b200/parallel.py). Array arguments are the FULL (Na, ...) arrays; the call works
 * on atoms [first, first+count). The caller reduces mesh / field / gradient arrays across ranks. */
int admp_set_box(admp_ctx* ctx, void* stream, const void* box);          /* cell + influence tables */
int admp_mesh_zero(admp_ctx* ctx, void* stream);
int admp_pme_spread_range(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols,
                          int M_stride, const void* U, int first, int count);
int admp_pme_gather_range(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols,
                          int M_stride, const void* U, int mode, uint32_t flags, void* dpos, void* G,
                          int G_stride, void* F, double* scalars, int first, int count);
int admp_pme_self_range(admp_ctx* ctx, void* stream, const void* M, const void* U, const void* pol,
                        uint32_t flags, void* G, void* F, void* dpol, double* scalars, int first,
                        int count);
int admp_frames_bwd_range(admp_ctx* ctx, void* stream, const void* pos, const void* Q_local,
                          const void* G, void* dQ_local, void* dpos, double* scalars, int first,
                          int count);
/* one optimize_Uind cycle on an assembled field (admp/pme.py:133-138): adds the self/penalty part to F,
 * tests max|F| < thresh over pol > 0.001 sites BEFORE updating U. state: device int32[8], zero before the
 * first cycle; state[0]=cycle, [3]=n_cycle, [4]=converged, [5]=continue. flags & ADMP_WANT_VIRIAL: the
 * caller runs its own final reciprocal pass (with the k-space virial sums) after the loop, so the loop
 * does not add a refresh pass after the last allowed update. */
int admp_scf_step(admp_ctx* ctx, void* stream, const void* M, void* U, const void* pol, void* F,
                  int maxiter, double thresh, uint32_t flags, int32_t* state, double* scalars);
/* folds the reciprocal-space accumulators of `scalars` into dE/dbox (ADMP_S_DBOX) */
int admp_virial_finalize(admp_ctx* ctx, void* stream, double* scalars);

/* energy_pme / get_energy / get_forces incl. optimize_Uind (admp/pme.py:58-143, :176-254).
 * U_io (n,3): in = U_init, out = converged U (ignored when non-polarizable: pass NULL with
 * pol/tholes/pScales NULL). scf_out[0]=n_cycle, [1]=converged flag (device int32[2]). */
int admp_pme_eval(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                  const int32_t* pairs, int64_t n_rows, const void* Q_local, void* U_io,
                  const void* pol, const void* tholes, const void* mScales, const void* pScales,
                  uint32_t flags, int maxiter, double thresh, double* scalars, void* dpos,
                  void* dQ_local, void* F, void* dpol, void* dtholes, int32_t* scf_out);

/* energy_disp_pme (admp/disp_pme.py:80-279): c_list (n,3) reals. dc (n,3) may be NULL. */
int admp_disp_eval(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                   const int32_t* pairs, int64_t n_rows, const void* c_list, const void* mScales,
                   int pmax, uint32_t flags, double* scalars, void* dpos, void* dc);

/* generate_pairwise_interaction(TT_damping_qq_c6_kernel) (admp/pairwise.py:45-113).
 * params: a, b, q, c (n each). dparams (4 pointers worth, (4,n) reals) may be NULL. */
int admp_tt_pair(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                 const int32_t* pairs, int64_t n_rows, const void* mScales, const void* a,
                 const void* b, const void* q, const void* c, uint32_t flags, double* scalars,
                 void* dpos, void* dparams);
/* the same pair pass with Tang-Toennies damped C8 and C10 terms added (sum_{n=6,8,10} e^-br P_n(br) c_n,i c_n,j / r^n; no
 * reference counterpart: SURVEY 8(f) rank 3). params: a, b, q, c6, c8, c10; dparams (6, n). */
int admp_tt_pair_c10(admp_ctx* ctx, void* stream, const void* pos, const void* box,
                     const int32_t* pairs, int64_t n_rows, const void* mScales, const void* a,
                     const void* b, const void* q, const void* c6, const void* c8, const void* c10,
                     uint32_t flags, double* scalars, void* dpos, void* dparams);

/* The generic half of generate_pairwise_interaction (admp/pairwise.py:57-77) for ANY user pair kernel: per row the
 * minimum-image distance dr (1.0 on rows that are not evaluated) and the scale index covalent_map[i,j]-1 with 0 -> 4
 * (-1 on rows that are not evaluated); and its adjoint: g_dr = dE/d(dr) per row -> dpos (n,3) (zeroed, then accumulated)
 * and, with ADMP_WANT_VIRIAL, scalars[ADMP_S_DBOX..] (zeroed, then the image-shift term). The user kernel runs in between
 * on device arrays (admp_b200/pairwise.py evaluates it with tensor operations; a JAX host would trace it). */
int admp_pair_geometry(admp_ctx* ctx, void* stream, const void* pos, const void* box, const int32_t* pairs,
                       int64_t n_rows, void* dr, int32_t* sidx);
int admp_pair_geometry_bwd(admp_ctx* ctx, void* stream, const void* pos, const void* box, const int32_t* pairs,
                           int64_t n_rows, const void* g_dr, uint32_t flags, void* dpos, double* scalars);

/* jax_md.partition.neighbor_list(..., format=OrderedSparse) replacement (call sites:
 * examples/water_1024/run_admp.py:109-112). pairs (capacity,2) int32 padded with (N,N);
 * info[0] = number of pairs found, info[1] = overflow flag (device int32[2]). */
int admp_nblist_build(admp_ctx* ctx, void* stream, const void* pos, const void* box, int n_atoms,
                      double rc, int32_t* pairs, int64_t capacity, int32_t* info);
/* same with the caller's host copy of the box (9 doubles): no device-to-host read of the box, no host synchronisation */
int admp_nblist_build_hostbox(admp_ctx* ctx, void* stream, const void* pos, const void* box, const double* box_host,
                              int n_atoms, double rc, int32_t* pairs, int64_t capacity, int32_t* info);

/* ---- x-slab decomposition of reciprocal space over the GPUs of one NVLink domain (no reference counterpart;
 * north-star config "256k-water box at 1/2/4/8 B200"). Rank r owns the x planes [floor(r*K1/n), floor((r+1)*K1/n)) of the mesh
 * and of the half spectrum; every rank maps the other ranks' buffers (cudaIpc) and the kernels address the owner
 * of each plane directly over NVLink: spread / gather for stencils crossing a slab boundary, and the fused X pass,
 * which reads and writes all ranks' planes (the transposes of a distributed FFT happen inside the kernel).
 * The caller orders the stages with stream-ordered cross-rank barriers:
 *   slab_zero | barrier | slab_spread | barrier | slab_fft 0 | barrier | slab_fft 1 | barrier | slab_fft 2 |
 *   barrier | slab_gather | (collective on the gathered field = barrier) */
int admp_ipc_export(const void* devptr, void* handle64);          /* cudaIpcGetMemHandle */
int admp_ipc_open(const void* handle64, void** devptr);           /* cudaIpcOpenMemHandle (peer access enabled) */
int admp_ipc_close(void* devptr);
/* mesh_ptrs / spec_ptrs: host arrays of nranks device pointers (entry `rank` = admp_ctx_buffer(ctx, 0 / 1)).
 * nranks = 0 clears the table. Needs K1 >= nranks and the register-blocked FFT kernels. */
int admp_ctx_set_peers(admp_ctx* ctx, int rank, int nranks, void* const* mesh_ptrs, void* const* spec_ptrs);
int admp_slab_zero(admp_ctx* ctx, void* stream);
int admp_slab_spread(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols, int M_stride,
                     const void* U, int count);
int admp_slab_fft(admp_ctx* ctx, void* stream, int phase, int kind, uint32_t flags, double* scalars);
int admp_slab_gather(admp_ctx* ctx, void* stream, const void* pos, const void* M, int M_cols, int M_stride,
                     const void* U, int mode, uint32_t flags, void* dpos, void* G, int G_stride, void* F,
                     double* scalars, int count);

/* Measurement helper (no reference counterpart): dense FMA throughput of the FP64 (dtype ADMP_F64) or FP32
 * CUDA-core pipe of the current device in TFLOP/s - the roofline denominator of the pair kernel. Allocates
 * and synchronises; not a compute entry point. */
int admp_fp_peak(void* stream, int dtype, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ADMP_B200_H */
