/*
 * fds_b200 -- C ABI of the B200-native finite-difference-shadowing hot path.
 *
 * One context per GPU (one process per GPU).  The context owns the
 * device-resident state of the shadowing segment loop of qiqi/fds:
 *
 *   ensemble  E[site][col]   site-major ("interleaved SoA"), col 0 = primal u,
 *                            cols 1..m = u + eps*W_j, col m+1 = u + eps*w
 *   tangents  T[site][j]     j < m homogeneous (W), j = m inhomogeneous (w)
 *   d[site]                  per-step du/dt estimate (time dilation)
 *
 * All host buffers are caller-owned row-major float64; all device memory is
 * library-owned.  Every function returns 0 on success, non-zero on failure;
 * fds_last_error() gives the message.  One host thread per context; no global
 * state; do not fork() after fds_ctx_create().
 *
 * Each entry point cites the reference (qiqi/fds) interface it replaces.
 */
#ifndef FDS_B200_H
#define FDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fds_ctx fds_ctx;

/* solver "systems" (the plugin on the far side of run(), fds/fds.py:184-200) */
enum {
    FDS_SYS_L96_RK4   = 0, /* Lorenz-96, classical RK4, dt, F = s + 8            */
    FDS_SYS_L96_EULER = 1, /* Lorenz-96, forward Euler: tests/solvers/mock_fun3d/fun3d:48-57 */
    /* small ODE systems of the reference test-suite (world = 1, n_local = state dimension);
     * their objective is recorded directly, not as sums over sites:                          */
    FDS_SYS_LORENZ63  = 2, /* n = 3, J = [(z-28)^2, 100]: tests/solvers/lorenz/lorenz.f03:10-26,
                              tests/test_lorenz.py:18-31                                     */
    FDS_SYS_VANDERPOL = 3, /* n = 2, J = [x0^2, -]: tests/solvers/vanderpol/vanderpol.f03:9-24 */
    FDS_SYS_CIRCULAR  = 4, /* n = 2, J = [x0^2+x1^2, -]: tests/solvers/circular/circular.f03:9-28 */
    /* no built-in solver: the caller's own run(u, s, steps) advances the trajectories on the host
     * (any n_local, any objective) and only the algebra of the segment loop runs here              */
    FDS_SYS_HOST      = 5
};

/* arithmetic modes */
enum {
    FDS_MATH_EXACT = 0,    /* FMA-free, numpy operation order: bit-identical state */
    FDS_MATH_FAST  = 1     /* FMA contraction allowed (statistically equivalent)   */
};

/* `build` argument of fds_boundary_finish / fds_boundary */
enum {
    FDS_BUILD_ENSEMBLE      = 1, /* write the next perturbed ensemble u + eps*[W; w]                */
    FDS_BUILD_KEEP_TANGENTS = 2  /* also store W, w explicitly (for fds_get_tangents right after);
                                    without it they are recovered from the ensemble, (E_k - E_0)/eps,
                                    to ~1e-9 relative -- enough to continue, not for a checkpoint   */
};

/* context flags */
enum {
    FDS_FLAG_SIMPLE_KERNELS  = 1, /* use the one-thread-per-element step kernel (cross-check path) */
    FDS_FLAG_EXTERNAL_STREAM = 2  /* launch on cfg.stream even if it is 0 (the legacy default stream);
                                     without it a NULL stream means "create my own"                 */
};

/* device buffers a multi-GPU host may need to address (halo exchange, reductions) */
enum {
    FDS_BUF_ENSEMBLE = 0,  /* current ensemble, base of site 0, [.][M]             */
    FDS_BUF_PRIMAL   = 1,  /* primal work array P0, base of site 0, [.]            */
    FDS_BUF_GRAM     = 2,  /* extended Gram (m+2)^2, to be sum-reduced over ranks  */
    FDS_BUF_JSUM     = 3,  /* objective sums [steps][M][2] of the last segment     */
    FDS_BUF_TANGENT  = 4,  /* explicit tangents [n_local][m+1]                     */
    FDS_BUF_DXDT     = 5   /* d[n_local]                                           */
};

typedef struct fds_config {
    int32_t device;        /* CUDA device ordinal                                  */
    int32_t system;        /* FDS_SYS_*                                            */
    int32_t math;          /* FDS_MATH_*                                           */
    int32_t m;             /* subspace dimension (number of homogeneous tangents)  */
    int64_t n_local;       /* sites owned by this rank                             */
    int64_t n_global;      /* sites of the whole ring                              */
    int64_t site_offset;   /* global index of local site 0                         */
    int32_t rank, world;   /* position in the ring of ranks                        */
    int32_t max_halo_steps;/* steps between halo refreshes (0 = library default)   */
    int32_t flags;         /* FDS_FLAG_*                                           */
    double  dt;            /* time step (0 = 0.01)                                 */
    void*   stream;        /* cudaStream_t to launch on (NULL = own stream)        */
    int32_t batch;         /* > 1: that many INDEPENDENT problems of n_local sites
                              each in one context (parameter sweep; world = 1,
                              Lorenz-96 RK4, n_local a multiple of 64)            */
    int32_t reserved;
} fds_config;

/* ---- life cycle ------------------------------------------------------------- */
int  fds_ctx_create(const fds_config* cfg, fds_ctx** out);
void fds_ctx_destroy(fds_ctx* ctx);
const char* fds_last_error(const fds_ctx* ctx);   /* ctx may be NULL: creation error */
const char* fds_version(void);

/* ---- state in / out (reference: Checkpoint.u0/V/v, fds/checkpoint.py:6) ------- */
int fds_set_state(fds_ctx*, const double* u_host);                 /* [n_local]      */
int fds_get_state(fds_ctx*, double* u_host);
int fds_set_tangents(fds_ctx*, const double* V_host, const double* v_host); /* [m][n_local], [n_local] */
int fds_get_tangents(fds_ctx*, double* V_host, double* v_host);
/* page-locked host memory (cudaHostAlloc) for the buffers above: optional -- any host pointer is
 * accepted -- but pinned buffers move at full PCIe rate and let uploads overlap the transposes  */
void* fds_host_alloc(size_t bytes);                                  /* NULL on failure */
void* fds_host_alloc_wc(size_t bytes);  /* write-combined: for buffers the host only writes (uploads) */
void  fds_host_free(void* p);
int fds_random_tangents(fds_ctx*, uint64_t seed);  /* uniform[0,1) V (pascal.random, fds/fds.py:23), v = 0 */

/* FD stencil weights of the time-dilation derivative (fds/timedilation.py:16-19 gets them
 * from np.linalg.solve; pass those 4 doubles for bit parity).  Default: -11/6, 3, -3/2, 1/3 */
int fds_set_dxdt_weights(fds_ctx*, const double c[4]);
/* any order of accuracy (TimeDilation.order_of_accuracy, fds/timedilation.py:8-12,64): `order` look-ahead steps
 * (1 .. 8, at most the halo capacity), order + 1 weights; Lorenz-96 device systems only for order != 3    */
int fds_set_dxdt_stencil(fds_ctx*, int order, const double* c);
int fds_dilation_order(const fds_ctx*);
/* how the time-dilation vector d (per-step du/dt) of a boundary is obtained:                */
enum {
    FDS_DILATION_OFF      = 0, /* run_ddt=0 of the reference (fds/timedilation.py:59-61)               */
    FDS_DILATION_FD       = 1, /* default: 3rd-order one-sided differences of 3 primal steps (:63-82)  */
    FDS_DILATION_ANALYTIC = 2, /* d = dt * f(u), the right-hand side of the built-in Lorenz-96 system   */
    FDS_DILATION_GIVEN    = 3  /* the caller's own run_ddt(u0, parameter) (TimeDilationExact, :54-61 --
                                  broken upstream, NameError at :58): upload it with fds_set_dxdt
                                  before every fds_boundary_dxdt                                       */
};
int fds_set_time_dilation(fds_ctx*, int mode);
int fds_set_dxdt(fds_ctx*, const double* d_host);   /* [n_local] (batch: [B][n_local]) */

/* thin QR of the tangent block (LssTangent.checkpoint, fds/lsstan.py:20-34): CholeskyQR from the
 * extended Gram.  passes = 0 (default): ONE pass when the first factor certifies that the loss of
 * orthogonality cond^2 u is below 2e-13 (the usual case: one segment of growth leaves the block
 * well conditioned), otherwise CholeskyQR2; passes = 2: always CholeskyQR2.
 * fds_last_condition: the bound sqrt(|G|_inf |G^-1|_inf) >= cond_2 of the last single-pass attempt. */
int    fds_set_qr_passes(fds_ctx*, int passes);
double fds_last_condition(const fds_ctx*);
/* split-phase API: after fds_boundary_factor, 1 if the certified single pass was accepted -- no second
 * Gram was launched, so the caller skips the second sum-reduce of FDS_BUF_GRAM                          */
int    fds_single_pass_pending(const fds_ctx*);
int    fds_last_qr_passes(const fds_ctx*);   /* 1, 2 or 3 (shifted CholeskyQR3): what the last boundary with QR took (after its results were read) */

/* ---- the solver plugin itself: u1, J = run(u0, s, steps), fds/fds.py:184-200 -- */
/* single trajectory on the context's primal array (world == 1; halos wrapped inside);
 * Jsum_host is [steps][2] = [sum x, sum x^2] (divide by n_global for J) or NULL      */
int fds_run(fds_ctx*, double s, int64_t steps, double* Jsum_host);
/* the same in chunks for world > 1: the host refreshes the primal halo (valid for
 * fds_halo_steps() steps) before each chunk; sums land in FDS_BUF_JSUM[step0+j][1][2] */
int fds_primal_advance(fds_ctx*, double s, int64_t step0, int64_t nsteps, int record);
/* tell the library the host has refreshed a halo itself (world > 1)                   */
int fds_halo_mark_valid(fds_ctx*, int which_buf, int64_t halo_steps);

/* ---- halo handling ----------------------------------------------------------- */
/* world == 1: periodic wrap on the device.  world > 1: the host exchanges the
 * regions described by fds_halo_layout() between ring neighbours (NCCL send/recv)
 * INSTEAD of calling these.  (reference: tests/solvers/mock_fun3d/fun3d:29-37)    */
int fds_halo_wrap_ensemble(fds_ctx*, int64_t halo_steps);
int fds_halo_wrap_primal(fds_ctx*, int64_t halo_steps);
/* element offsets (in doubles, relative to the FDS_BUF_* base) and counts of the
 * four contiguous regions: send_to_right, send_to_left, recv_from_left, recv_from_right */
int fds_halo_layout(const fds_ctx*, int which_buf, int64_t halo_steps, int64_t off[4], int64_t cnt[4]);

/* ---- one segment boundary, split at its reduction points ---------------------- */
/* reference: TimeDilation.__init__/contribution/project (fds/timedilation.py:37-82),
 * LssTangent.checkpoint (fds/lsstan.py:20-34), perturbed ICs (fds/segment.py:43-51).
 *
 *  fds_boundary_begin   : primal <- ensemble col 0 (or the state set by fds_set_state)
 *        [halo: primal, 3 steps]
 *  fds_boundary_dxdt    : u1,u2,u3, d = sum c_i u_i (or the analytic / given d, fds_set_time_dilation);
 *                         Gram of [tangents ; d] -> FDS_BUF_GRAM
 *        [sum-reduce FDS_BUF_GRAM over ranks]
 *  fds_boundary_factor  : G_dil, g_dil, projection.  do_qr: if the Gram certifies the block well
 *                         conditioned, the single-pass CholeskyQR factor is kept for _finish and nothing
 *                         else happens here; otherwise CholeskyQR pass 1 is applied and the Gram of the
 *                         result goes to FDS_BUF_GRAM
 *        [if do_qr: sum-reduce FDS_BUF_GRAM over ranks (harmless after a single-pass factor)]
 *  fds_boundary_finish  : R, b (single pass, or CholeskyQR pass 2); explicit tangents; next ensemble
 *        [halo: ensemble, halo_steps]
 */
int fds_boundary_begin(fds_ctx*, int from_ensemble);
int fds_boundary_dxdt(fds_ctx*, double s, int from_ensemble, double eps);
int fds_boundary_factor(fds_ctx*, int do_qr, int from_ensemble, double eps);
int fds_boundary_finish(fds_ctx*, int do_qr, double eps, int build_ensemble);
/* small results of the last boundary: R[m][m], b[m], G_dil[m], g_dil[1]; any may be NULL */
int fds_boundary_results(fds_ctx*, double* R, double* b, double* G_dil, double* g_dil);

/* ---- one segment: advance all m+2 trajectories `steps` steps ------------------ */
/* reference: run_segment's m+2 solver calls (fds/segment.py:56-64).  The host
 * refreshes halos every fds_halo_steps() steps: call fds_segment_advance() in
 * chunks of at most that many steps (step0 = index of the first step of the chunk
 * within the segment).  Inside a chunk the Lorenz-96 RK4 system advances TWO time
 * steps per kernel pass (temporal blocking; bit-identical to single steps).        */
int64_t fds_halo_steps(const fds_ctx*, int64_t steps);
int fds_segment_advance(fds_ctx*, double s, double eps, int64_t step0, int64_t nsteps);
/* objective sums of the last segment, J_host[steps][m+2][2] = [sum x, sum x^2]    */
int fds_segment_jsums(fds_ctx*, int64_t steps, double* Jsum_host);

/* ---- whole boundary / whole segment in one call (all phases, halos and reductions inside) -----
 * world == 1, or world > 1 with a communicator attached (fds_comm_init below).                  */
int fds_boundary(fds_ctx*, double s, double eps, int from_ensemble, int do_qr, int build_ensemble,
                 int64_t halo_steps, double* R, double* b, double* G_dil, double* g_dil);
int fds_segment(fds_ctx*, double s, double eps, int64_t steps, double* Jsum_host);
/* The same WITHOUT waiting for the device: everything one iteration of the reference's loop does
 * (fds/fds.py:128-170) is put on the context's stream and the call returns.  For m = 16 / 32 the adaptive
 * CholeskyQR decision (one pass or two) is taken on the device, so a boundary has no host round trip at
 * all; the results land in caller-owned (preferably page-locked, fds_host_alloc) memory when the stream
 * gets there:
 *   out[fds_boundary_out_doubles()] per problem = R[m*m], b[m], G_dil[m], g_dil,
 *                                                 {int32 status, int32 two_passes_taken}, cond bound
 *   Jsum_host[steps][m+2][2]
 * Call fds_sync(), then fds_boundary_check(out, do_qr) (non-zero = a Cholesky pivot failed, message in
 * fds_last_error).  out / Jsum_host may be NULL.                                       */
int64_t fds_boundary_out_doubles(const fds_ctx*);
int fds_boundary_enqueue(fds_ctx*, double s, double eps, int from_ensemble, int do_qr, int build_ensemble,
                         int64_t halo_steps, double* out);
int fds_boundary_check(fds_ctx*, const double* out, int do_qr);
int fds_segment_enqueue(fds_ctx*, double s, double eps, int64_t steps, double* Jsum_host);

/* ---- multi-GPU inside the library: NCCL on the context's stream ---------------------------------
 * The reference shards the state over MPI ranks (fds/compute.py:87-89) and exchanges halos / reduces
 * dot products over MPI (tests/solvers/mock_fun3d/fun3d:29-37, pascal_lite/operators/plinalg.py:19-64).
 * Here rank 0 obtains a 128-byte id (fds_comm_unique_id), the host distributes it by whatever means it
 * has (MPI_Bcast, torch.distributed, a file), and every rank calls fds_comm_init on its context (created
 * with cfg.rank / cfg.world).  From then on fds_halo_wrap_*, fds_run, fds_boundary*, fds_segment* work for
 * world > 1: ring send/recv of the halos, one all-reduce of the (m+2)^2 Gram per CholeskyQR pass, and an
 * all-gather + FIXED-order tree sum of the objective partials (J is bitwise independent of the number of
 * ranks for power-of-two shard sizes).  NCCL is bound at run time (dlopen of libnccl.so.2, or $FDS_NCCL_LIB).
 * fds_comm_attach: use an existing ncclComm_t (not owned) instead.                                  */
int fds_comm_unique_id(void* id128);
int fds_comm_init(fds_ctx*, const void* id128);
int fds_comm_attach(fds_ctx*, void* nccl_comm);
int fds_comm_attached(const fds_ctx*);
/* host-buffer form of one reference run_segment call (fds/segment.py:14-82):
 * u, V, v in -> advanced u, V, v out, Jsum[steps][m+2][2]; H2D and D2H inside     */
int fds_run_segment_host(fds_ctx*, double* u, double* V, double* v, double s, double eps,
                         int64_t steps, double* Jsum_host);

/* ---- host-solver mode (cfg.system = FDS_SYS_HOST) -------------------------------------------
 * For solvers that cannot move to the GPU (the reference's CFD applications, apps/): the perturbed
 * initial conditions u0 + eps*[V; v] come out as E_host[m+2][n] (row 0 = u0, rows 1..m homogeneous,
 * row m+1 inhomogeneous -- the inputs of the m+2 run() calls of fds/segment.py:56-64), the advanced
 * states go back in, and time dilation takes the states after 1, 2, 3 steps
 * (fds/timedilation.py:63-82).  Everything else (Gram, CholeskyQR, projection, R, b, G_dil) is
 * fds_boundary_* as usual.  fds_run / fds_segment_advance fail for this system.                 */
int fds_get_ensemble(fds_ctx*, double* E_host);                    /* [m+2][n] */
int fds_set_ensemble(fds_ctx*, const double* E_host, double eps);  /* [m+2][n] */
int fds_set_primal_history(fds_ctx*, int j, const double* u_host); /* j = 1, 2, 3 */

/* ---- batched parameter sweep (BASELINE configs[4]) -----------------------------
 * A context created with cfg.batch = B holds B independent problems; the reference can only
 * loop over them (tests/test_lorenz.py:35-41, apps/lorenz_demo.py:86-89).  Every call above acts
 * on all problems at once; host arrays gain a leading problem axis -- u[B][n], V[B][m][n],
 * v[B][n], R[B][m][m], b[B][m], G_dil[B][m], g_dil[B], Jsum[steps][B][m+2][2] -- and problem p
 * runs at parameter s + ds[p].  fds_segment_advance takes at most fds_halo_steps() per call.   */
int fds_set_batch_params(fds_ctx*, const double* ds);   /* ds[B] (NULL = all zero) */
/* what run_segment returns besides the state (fds/segment.py:75,79), reduced on the device so that
 * only KBs per problem cross PCIe: J0[B][steps][2] = objective history of the primal (means over
 * sites), G[B][m+1][2] = trapez_mean(J_j - J_0) / eps, rows 0..m-1 homogeneous, row m inhomogeneous.
 * Works for B = 1 too.                                                                          */
int fds_segment_sensitivities(fds_ctx*, int64_t steps, double eps, double* J0_host, double* G_host);
int fds_batch_count(const fds_ctx*);
int fds_lss_solve_batch(int device, const double* R, const double* b, int64_t K, int m, int64_t nbatch,
                        double* alpha);                  /* R[B][K][m][m], b / alpha [B][K][m] */

/* ---- LssTangent.solve (fds/lsstan.py:36-50): min |a|^2 s.t. a_{i+1}=R_i a_i+b_i */
/* R[K][m][m], b[K][m] -> alpha[K][m]; single-CTA block-tridiagonal Cholesky       */
int fds_lss_solve(int device, const double* R, const double* b, int64_t K, int m, double* alpha);

/* ---- introspection ------------------------------------------------------------ */
void*   fds_buffer_ptr(const fds_ctx*, int which_buf);   /* device pointer of site/elem 0 */
int64_t fds_kernel_launches(const fds_ctx*);              /* kernels launched so far        */
int64_t fds_collectives(const fds_ctx*);                  /* NCCL operations enqueued so far */
/* a slab of the current ensemble, rows [site0, site0 + nsites) of E[site][m+2] (ghost sites allowed), and
 * one whole column k (0 = primal, fds/segment.py:43; m+1 = the parameter-perturbed run, :64): what a test
 * needs to compare any window of a large state with the reference's run() on the same window            */
int fds_get_ensemble_rows(fds_ctx*, int64_t site0, int64_t nsites, double* out);   /* [nsites][m+2] */
int fds_get_ensemble_column(fds_ctx*, int k, double* out);                         /* [n_local]     */
/* strip decomposition of the step kernel: out = {a0, strip length, strips, 8-blocks per strip, log2 sites
 * per objective node, nodes, lanes a CTA packs (strip, column) pairs over (0 = whole strips), steps per pass} */
int fds_strip_plan(const fds_ctx*, int64_t out[8]);
/* measured roof of the exact-mode step kernel: FP64 warp-instructions per second of the whole GPU with the
 * kernel's own shape and instruction mix (~0.25 s of load, so at the clock the board sustains)          */
int fds_fp64_issue_rate(int device, double* winst_per_s, double* seconds, double* winst_per_clk_sm);
int     fds_sync(fds_ctx*);
/* per-launch device timing of the ensemble step kernel (CUDA events on the context's stream):
 * enable, run, then read the summed duration and the number of launches since enabling   */
int     fds_profile_enable(fds_ctx*, int enabled);
int     fds_profile_read(fds_ctx*, double* total_ms, int64_t* launches);
int     fds_profile_cycles(fds_ctx*, double* mean_cycles_per_cta);   /* SM cycles a CTA of those launches ran, clock64() inside the kernel */
int64_t fds_profile_steps(const fds_ctx*);   /* time steps those launches covered (2 per launch with temporal blocking) */

#ifdef __cplusplus
}
#endif
#endif /* FDS_B200_H */
