// Internal launch API of the fds_b200 CUDA kernels (C++; the C ABI is in ctx.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "l96_core.cuh"

namespace fds {

struct LaunchStats { long long launches = 0; long long collectives = 0; };

// ---- l96_step.cu ------------------------------------------------------------ //
struct StepArgs {
    const double* X;       // input, pointer to site 0, layout [site][M]
    double* Y;             // output, same layout
    int M;                 // columns
    long long n;           // owned sites (objective sums cover [0, n))
    long long lo, hi;      // output range of this step (simple kernels only; the streaming
                           // kernel always computes the whole span of its fixed StripPlan)
    double F0, Flast;      // forcing of columns < M-1 and of column M-1
    double dt;
    int exact;             // FDS_MATH_EXACT / FAST
    double* jpart;         // simple kernels: [nleaf][M][2] 64-site partial sums; streaming kernel:
                           // [plan.nnode][M][2] node partial sums; or nullptr
    long long nleaf;
    // batch of independent problems stacked along the site axis, `slot` sites apart (0 = no batch):
    // forcing pair (F0, Flast) of problem p in Fpp[p]; the streaming kernel cuts every slot into
    // `spp` strips, the simple kernel maps site i to problem i / slot
    const double2* Fpp = nullptr;
    int spp = 0;
    long long slot = 0;
    // streaming kernel launch shape (0 = default: as many strips per CTA as fit 480 threads, one
    // CTA per SM with a dedicated producer warp).  ctas_per_sm = 2: <= 256 threads and half the
    // shared memory per CTA, warp 0 produces in between its blocks
    int runs_per_cta = 0;
    int ctas_per_sm = 1;
    // two RK4 steps in one pass (l96_march2.cuh): Y receives the state after TWO steps, jpart the node
    // partials of the first and jpart2 those of the second step; needs runs_per_cta * M <= 256 threads and the in-between producer
    int two_steps = 0;
    double* jpart2 = nullptr;
    // K time steps in one pass (2: as two_steps; 4: forward Euler only, l96_marchk.cuh): step j's node partials go to
    // jpart + j * nnode * M * 2
    int nfuse = 0;
    // RK4, two steps per pass: also write column 0 of the output (the primal state) to this single-column array
    // (pointer to site 0) -- the last pass of a segment saves the strided extraction pass of the next boundary
    double* primal_out = nullptr;
    // profiling: every CTA adds its clock64() span to cycles[0] and 1 to cycles[1] (nullptr: off)
    unsigned long long* cycles = nullptr;
};

// streaming TMA-fed kernel (RK4 only); `plan` = make_strip_plan(n, H, num_sms * stream_runs_per_cta(M)),
// X and Y allocated for sites [plan.alo, plan.ahi)
int stream_runs_per_cta(int M);
int two_step_runs_per_cta(int M);      // strips per CTA of the two-steps-per-pass kernel (256 threads)
// euler = 1: forward Euler with the reference's operation order instead of RK4
cudaError_t launch_l96_rk4_stream(const StepArgs& a, const StripPlan& plan, cudaStream_t st, LaunchStats* ls,
                                  int euler = 0);
// one thread per element, global loads (RK4 or Euler); no objective sums
cudaError_t launch_l96_simple(const StepArgs& a, int euler, cudaStream_t st, LaunchStats* ls);
// d[i] = dt * f(X)_i for i in [0, n) of a single-column state with valid ghosts (2 left, 1 right)
cudaError_t launch_l96_ddt(const double* X, double* d, long long n, double F0, double dt, int exact, int euler,
                           const double2* Fpp, long long slot, cudaStream_t st, LaunchStats* ls);
// 64-site leaf partials of an array: jpart[leaf][M][2]
cudaError_t launch_stats_leaf(const double* Y, int M, long long n, double* jpart, long long nleaf,
                              cudaStream_t st, LaunchStats* ls);
// canonical tree over the leading axis of in[n_in][C] -> out[C]; scratch holds 2 * ceil(n_in/256) * C
cudaError_t launch_tree_reduce(const double* in, long long n_in, int C, double* out, double* scratch,
                               cudaStream_t st, LaunchStats* ls);
// batch of independent reductions: in[batch][n_in][C] -> out[batch][C]; scratch: 2 * batch * ceil(n_in/256) * C
cudaError_t launch_tree_reduce_batched(const double* in, long long n_in, int C, long long batch, double* out,
                                       double* scratch, cudaStream_t st, LaunchStats* ls);
// periodic ghosts: A[g][:] = A[g mod n][:] for g in [-gl,0) U [n, n+gr)
cudaError_t launch_halo_wrap(double* A, int M, long long n, long long gl, long long gr,
                             cudaStream_t st, LaunchStats* ls);
// the same for `batch` problems whose site 0 is `slot` sites apart (A points at site 0 of problem 0)
cudaError_t launch_halo_wrap_batch(double* A, int M, long long n, long long gl, long long gr, int batch,
                                   long long slot, cudaStream_t st, LaunchStats* ls);
// objective sums of a batch: leaves[step][nleaf_total][C] (64-site partials of the stacked array) ->
// out[step][p][C], canonical tree over the n/64 leaves of problem p (first leaf: leaf0 + p*slot_leaves)
cudaError_t launch_batch_jreduce(const double* leaves, long long nleaf_total, long long leaf0, long long slot_leaves,
                                 long long leaves_pp, int C, int batch, long long steps, double* out,
                                 cudaStream_t st, LaunchStats* ls);

// cross-rank continuation of the canonical tree: gathered[world][len] (rank-major) -> out[len], balanced binary
// tree over the ranks, zero padded to a power of two, lower rank on the left (fds_b200/ring.py tree_sum_leading)
cudaError_t launch_rank_tree_sum(const double* gathered, int world, long long len, double* out, cudaStream_t st,
                                 LaunchStats* ls);

// batch: J0[p][steps][2] (means) and G[p][m+1][2] (trapezoid-mean FD sensitivities / eps) from jsum[steps][nb][M][2]
cudaError_t launch_batch_sens(const double* jsum, int nb, int M, long long steps, double n_sites, double eps,
                              double* J0, double* G, cudaStream_t st, LaunchStats* ls);

// ---- boundary.cu ------------------------------------------------------------ //
cudaError_t launch_extract_col0(const double* E, int M, double* P, long long n, cudaStream_t, LaunchStats*);
cudaError_t launch_dxdt(const double* P0, const double* P1, const double* P2, const double* P3,
                        const double* c4_dev, double* d, long long n, cudaStream_t, LaunchStats*);
// Time dilation in ONE pass (fds/timedilation.py:14-35,63-82): the `order` look-ahead states u_1 .. u_order of the
// primal (each one solver step, same arithmetic as launch_l96_simple) are formed tile by tile in shared memory and
// d = sum_k c_k u_k is accumulated in the reference's left-to-right order -- no intermediate state touches HBM.
// P0: site 0 of the single-column primal with ghosts valid for `order` steps; d: site 0.  Batch: stacked array + Fpp.
constexpr int kMaxDilationOrder = 8;
struct DilationStencil { int order; double c[kMaxDilationOrder + 1]; };
cudaError_t launch_dilation_fused(const double* P0, double* d, long long n, const DilationStencil& st, double F0,
                                  double dt, int exact, int euler, const double2* Fpp, long long slot,
                                  cudaStream_t, LaunchStats*);
cudaError_t launch_primal_multi(const double* X, double* Y, long long lo, long long hi, int steps, double F0, double dt,
                                int exact, int euler, double* jp, long long sites, long long nleaf,
                                const double2* Fpp, long long slot, cudaStream_t, LaunchStats*);
cudaError_t launch_diff_tangents(const double* E, int M, double eps, double* T, long long n,
                                 cudaStream_t, LaunchStats*);
cudaError_t launch_make_ensemble(const double* u, const double* T, int M, double eps, double* E,
                                 long long n, cudaStream_t, LaunchStats*);
// in[rows][cols] -> out[cols][ld_out] (first `rows` entries of each output row)
cudaError_t launch_transpose(const double* in, long long rows, int cols, double* out, long long ld_out,
                             cudaStream_t, LaunchStats*);
// in[cols][ld_in] -> out[rows][cols]
cudaError_t launch_transpose_back(const double* in, long long ld_in, long long rows, int cols, double* out,
                                  cudaStream_t, LaunchStats*);
cudaError_t launch_fill_random(double* T, long long n, int MO, int m, long long site_offset,
                               long long n_global, uint64_t seed, cudaStream_t, LaunchStats*);

enum { SRC_ENSEMBLE = 0, SRC_TANGENT = 1 };

// Batch of independent problems for the boundary kernels: problem p = blockIdx.y works on arrays
// offset by p * slot sites; its small matrices live p * small_stride doubles further on.
struct BatchDesc {
    int count = 1;
    long long slot = 0;          // sites between consecutive problems
    long long small_stride = 0;  // doubles between consecutive SmallBufs blocks
    long long part_stride = 0;   // doubles between consecutive Gram partial blocks
    // device-side gate: when set, the kernel returns at once unless *gate == gate_want.  The adaptive
    // CholeskyQR decision (one pass or two) is taken ON the device by small_single_kernel, so both
    // variants of a boundary are enqueued without a host round trip and the untaken one costs an
    // empty launch.
    const int* gate = nullptr;
    int gate_want = 0;
    long long gate_stride = 0;   // ints between the gates of consecutive problems of a batch (0: one gate for all)
};
// Gram of X_i = [t_0..t_m, d_i]  (MX = m+2), t from the ensemble ((E_k - E_0) * inv_eps) or from T.
// part: [grid][MX*MX] scratch; out: MX*MX (deterministic fixed-order sum of the partials)
cudaError_t launch_gram(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                        long long n, double* part, int max_part, double* out, int num_sms,
                        cudaStream_t, LaunchStats*, const BatchDesc* batch = nullptr);
// out_i = A . X_i  (A: (m+1) x (m+2) row-major on device); writes T (may alias the source T) and,
// if E_out != nullptr, the next ensemble E_out[i] = [u_i, u_i + eps*out_i]
cudaError_t launch_apply(int src, const double* E, const double* Tin, const double* d, const double* A,
                         int m, double inv_eps, long long n, double* Tout, double* E_out, const double* u,
                         double eps, cudaStream_t, LaunchStats*, const BatchDesc* batch = nullptr);

struct SmallBufs {          // all device pointers
    double* gram;           // (m+2)^2
    double* A;              // (m+1)(m+2)
    double* L1;             // m*m
    double* R;              // m*m
    double* b;              // m
    double* gdil;           // m+1  (G_dil[0..m-1], g_dil)
    int* status;            // 0 ok, >0 Cholesky pivot failure (1-based column)
};
// status word: 0 ok; 1000 + column: Cholesky pivot <= 0 in pass 1; 2000 + column: in pass 2 (kSmallPass*).
// small2 only ever RAISES the status (a pass-1 failure stays visible when both are enqueued back to back).
// `gram2`: small2 factors this Gram instead of B.gram (the gated boundary keeps the two Grams apart).
constexpr int kSmallPass1 = 1000, kSmallPass2 = 2000;
cudaError_t launch_small1(const SmallBufs&, int m, int use_d, int do_qr, cudaStream_t, LaunchStats*,
                          const BatchDesc* batch = nullptr, double shift_rel = 0.0);
// middle pass of the shifted CholeskyQR3 (after a small1 with shift_rel > 0): Gram of [Q1; p_w] -> L2, L1 <- L1 L2, A
cudaError_t launch_small_mid(const SmallBufs&, int m, cudaStream_t, LaunchStats*, const BatchDesc* batch = nullptr);
cudaError_t launch_small2(const SmallBufs&, int m, cudaStream_t, LaunchStats*, const BatchDesc* batch = nullptr,
                          const double* gram2 = nullptr);
// single-pass CholeskyQR from one Gram; writes sqrt(|G|_inf |G^-1|_inf) >= cond_2 to ((double*)status)[1].
// two_pass_flag (optional): problem p's flag, two_pass_flag[p * flag_stride], is set to 1 when its factor failed or
// cond^2 u >= kSinglePassTol, or when force_two is set, else to 0 -- the gate of the CholeskyQR2 kernels enqueued
// behind it, problem by problem (a well-conditioned problem of a sweep does not pay for an ill-conditioned one).
constexpr double kSinglePassTol = 2e-13;
cudaError_t launch_small_single(const SmallBufs&, int m, int use_d, cudaStream_t, LaunchStats*,
                                const BatchDesc* batch = nullptr, int* two_pass_flag = nullptr, int force_two = 0,
                                long long flag_stride = 0);
size_t small_smem_bytes(int m);

// ---- boundary_fast.cu ---------------------------------------------------------- //
// register-blocked versions of launch_gram / launch_apply; A_host is the (m+1) x (m+2) matrix on
// the HOST (it travels as a by-value kernel parameter); Tout / E_out may be nullptr
cudaError_t launch_gram_fast(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                             long long n, double* part, int max_part, double* out, int num_sms, cudaStream_t,
                             LaunchStats*);
bool apply_fast_supported(int m);
cudaError_t launch_apply_fast(int src, const double* E, const double* Tin, const double* d, const double* A_host,
                              int m, double inv_eps, long long n, double* Tout, double* E_out, const double* u,
                              double eps, int num_sms, cudaStream_t, LaunchStats*);

// ---- boundary_mma.cu ----------------------------------------------------------- //
// Gram on the FP64 tensor cores (mma.sync m8n8k4); same contract as launch_gram_fast
bool gram_mma_supported(int m);
cudaError_t launch_gram_mma(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                            long long n, double* part, int max_part, double* out, int num_sms, cudaStream_t,
                            LaunchStats*, const BatchDesc* batch = nullptr);

// out_i = A X_i on the FP64 tensor cores; A in DEVICE memory; same contract as launch_apply_fast otherwise
bool apply_mma_supported(int m);
cudaError_t launch_apply_mma(int src, const double* E, const double* Tin, const double* d, const double* A, int m,
                             double inv_eps, long long n, double* Tout, double* E_out, const double* u, double eps,
                             int num_sms, cudaStream_t, LaunchStats*, const BatchDesc* batch = nullptr);

// ---- diag.cu ----------------------------------------------------------------- //
// FP64 issue-rate microbenchmark: the roof of the exact-mode step kernel.  Runs the step kernel's own
// instruction mix (2 DADD : 1 DMUL, 8 independent chains per thread, 2 warps per scheduler, one CTA per SM)
// and returns FP64 warp-instructions per second for the whole GPU at the clock the board sustains.
cudaError_t fp64_issue_rate(int num_sms, cudaStream_t st, double* winst_per_s, double* seconds, double* winst_per_clk_sm);

// ---- toy_ode.cu -------------------------------------------------------------- //
// small ODE systems (Lorenz-63, Van der Pol, circular): state dimension, 0 for the L96 systems
int toy_dim(int system);
// X, Y: [dim][M] (may alias); all nsteps in one launch; jout: [nsteps][M][2] objective values or nullptr
cudaError_t launch_toy_advance(int system, int exact, const double* X, double* Y, int M, long long nsteps,
                               double s0, double s_last, double* jout, cudaStream_t st, LaunchStats* ls);

// ---- lss.cu ----------------------------------------------------------------- //
// R[K][m][m], b[K][m] (device) -> alpha[K][m] (device); work: (2*m*m + 2*m) * K doubles
// batch > 1: R[batch][K][m][m], b / alpha [batch][K][m], work and status per problem (one CTA each)
cudaError_t launch_lss_solve(const double* R, const double* b, long long K, int m, double* alpha,
                             double* work, int* status, cudaStream_t, LaunchStats*, int batch = 1);
size_t lss_smem_bytes(int m);

}  // namespace fds
