// Context + C ABI of libfds_b200 (see include/fds_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fds_b200.h"
#include "kernels.h"
#include "nccl_dl.h"

using namespace fds;

struct fds_ctx {
    fds_config cfg;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int m = 0, M = 0, MO = 0, MX = 0;
    long long n = 0;
    long long H = 1;               // halo capacity in steps
    long long alo = 0, ahi = 0;    // allocated site range of ensemble / primal arrays
    StripPlan plan{};              // fixed strip decomposition of the ensemble step kernel
    bool stream_ok = false;
    bool fuse2 = false;            // two time steps per pass of the streaming kernel (l96_march2.cuh)
    bool fuse4 = false;            // forward Euler: four (l96_marchk.cuh)
    long long prof_steps = 0;      // time steps covered by the profiled launches
    bool toy = false;              // small ODE system (toy_ode.cu) or host solver: no stencil, no halos
    bool host_sys = false;         // FDS_SYS_HOST: the solver runs on the host, only the algebra lives here
    // batch of nb independent problems stacked along the site axis (cfg.batch > 1): problem p owns the
    // stacked sites [gl0 + p*slot, gl0 + p*slot + n); the gaps in between hold its ghost cells
    int nb = 1;
    long long slot = 0, gl0 = 0;
    long long nall = 0;            // stacked sites (= n without a batch)
    int spp = 1, runs_per_cta = 0, ctas_per_sm = 2; // streaming-kernel shape of the batch
    BatchDesc bd{};
    std::vector<double> ds;        // parameter offset of every problem
    std::vector<double2> Fhost;
    double2* Fpp = nullptr;        // device (F0, Flast) per problem
    double* sens = nullptr;        // device staging of fds_segment_sensitivities
    size_t sens_cap = 0;
    double* T_alloc = nullptr;     // stacked site 0 of the tangent array (T = site 0 of problem 0)
    double* d_alloc = nullptr;
    long long small_stride = 0;
    double* Ebase[2] = {nullptr, nullptr};
    int cur = 0;
    double* Pbase[4] = {nullptr, nullptr, nullptr, nullptr};
    double* T = nullptr;
    double* d = nullptr;
    double* gram_part = nullptr;
    int max_part = 0;
    double* small_block = nullptr;
    SmallBufs small{};
    double* c4 = nullptr;
    DilationStencil stencil{3, {-11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0, 0, 0, 0, 0, 0}};   // fds/timedilation.py:64: order 3
    double* A_host = nullptr;      // pinned copy of small.A (+ status): A travels to apply_fast as a kernel parameter
    bool fast_boundary = false;    // register-blocked gram / apply kernels (boundary_fast.cu)
    bool mma_boundary = false;     // tensor-core gram / apply kernels (boundary_mma.cu): single problem or batch
    double* jpart = nullptr;
    double* jscratch = nullptr;
    double* jsum = nullptr;
    long long jsum_cap = 0;        // steps
    double* jnode = nullptr;       // [H][plan.nnode][M][2] node partials of one halo chunk of the streaming kernel
    double* jnode_scratch = nullptr;
    long long nleaf = 0;
    // host <-> device staging of the tangent block: two buffers, a copy stream and events so
    // that the PCIe transfer of chunk i+1 overlaps the transpose of chunk i
    double* stage[2] = {nullptr, nullptr};
    long long stage_sites = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // state machine
    bool state_in_ensemble = false;   // newest primal state is column 0 of the ensemble
    bool tangents_implicit = false;   // newest tangents are (E_k - E_0)/eps_last
    bool primal_synced = false;       // ... and P(0) already holds that column (the last pass of the segment wrote it)
    bool time_dilation = true;
    int dil_mode = 1;                 // FDS_DILATION_*: 0 off, 1 finite differences, 2 analytic (dt * f(u)), 3 given by the host
    // CholeskyQR: 0 = one pass when the first factor certifies cond^2 u negligible, else two; 2 = always two
    int qr_mode = 0;
    bool qr_single = false;           // a certified single-pass factor is waiting for fds_boundary_finish
    int pend_src = 0;
    double pend_inv_eps = 0.0;
    double cond_last = 0.0;           // upper bound of cond_2 from the last single-pass attempt
    int passes_last = 0;              // CholeskyQR passes the last boundary with QR took (1, 2, or 3: shifted CholeskyQR3)
    bool shifted_last = false;
    double eps_last = 0.0;
    long long ens_halo_valid = 0, pri_halo_valid = 0;
    LaunchStats ls;
    std::string err;
    // in-library communicator (fds_comm_*): NCCL on the context's stream -- halo send/recv, Gram
    // all-reduce, all-gather + fixed-order sum of the objective partials; nothing returns to the host
    const NcclApi* nccl = nullptr;
    NcclComm comm = nullptr;
    bool own_comm = false;
    double* jgather = nullptr;        // [world][len] gathered objective partials
    size_t jgather_cap = 0;
    // device-gated boundary (adaptive CholeskyQR decided on the device, results fetched asynchronously)
    bool gated = false;
    double* gram2 = nullptr;          // Gram of the CholeskyQR2 second pass (kept apart from small.gram)
    double* out_host = nullptr;       // pinned mirror of [R, b, G_dil, g_dil, status, cond] of the last boundary
    // optional per-launch timing of the step kernel (fds_profile_*)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    unsigned long long* prof_cycles = nullptr;   // device: {sum of the CTAs' clock spans, CTAs} of the profiled launches

    // *_all: stacked site 0; the others: site 0 of problem 0 (the same thing without a batch)
    double* E_all(int which) const { return Ebase[which] + (-alo) * M; }
    double* E(int which) const { return E_all(which) + gl0 * M; }
    double* Ecur() const { return E(cur); }
    double* P_all(int j) const { return Pbase[j] + (-alo); }
    double* P(int j) const { return P_all(j) + gl0; }
    const BatchDesc* batch() const { return nb > 1 ? &bd : nullptr; }
};

// CholeskyQR2 needs cond_2 of the block below ~ u^-1/2 / poly(m, n); past this bound the shifted CholeskyQR3 runs
constexpr double kShiftedQrCond = 1e6;

static thread_local std::string g_create_err;   // creation / context-free errors, per calling thread

#define FDS_CK(ctx, call)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

static int fail(fds_ctx* c, const std::string& msg) { c->err = msg; return 2; }

// after small1 / small2: bring A (for the by-value kernel parameter) and the Cholesky status to the host
static int check_small_status(fds_ctx* c, const char* what) {
    int st = 0;
    if (c->nb > 1) {
        struct SC { int st; int pad; double cond; };
        std::vector<SC> all(c->nb, SC{0, 0, 0.0});
        FDS_CK(c, cudaMemcpy2DAsync(all.data(), sizeof(SC), c->small.status, (size_t)c->small_stride * sizeof(double),
                                    sizeof(SC), c->nb, cudaMemcpyDeviceToHost, c->stream));
        FDS_CK(c, cudaStreamSynchronize(c->stream));
        c->cond_last = 0.0;
        for (int p = 0; p < c->nb; ++p) {
            if (all[p].st != 0) return fail(c, std::string(what) + ": Cholesky pivot <= 0 at column " + std::to_string(all[p].st % 1000) +
                                               " of problem " + std::to_string(p) + " (tangent block numerically rank deficient)");
            c->cond_last = std::max(c->cond_last, all[p].cond);      // the worst problem decides
        }
        return 0;
    }
    struct { int st; int pad; double cond; } sc = {0, 0, 0.0};
    FDS_CK(c, cudaMemcpyAsync(&sc, c->small.status, sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
    if (c->fast_boundary)
        FDS_CK(c, cudaMemcpyAsync(c->A_host, c->small.A, (size_t)c->MO * c->MX * sizeof(double),
                                  cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    st = sc.st;
    c->cond_last = sc.cond;
    if (st != 0) return fail(c, std::string(what) + ": Cholesky pivot <= 0 at column " + std::to_string(st % 1000) +
                                    " (tangent block numerically rank deficient)");
    return 0;
}

// RAII: make the context's GPU current for the duration of an entry point and restore the caller's
// afterwards (a process may hold contexts on several GPUs, and torch may have changed the current device)
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define FDS_ON_DEVICE(c) DeviceGuard guard__((c)->cfg.device)

static cudaError_t do_gram(fds_ctx* c, int src, double inv_eps) {
    if (c->mma_boundary || (c->fast_boundary && gram_mma_supported(c->m)))
        return launch_gram_mma(src, c->Ecur(), c->T, c->d, c->m, inv_eps, c->n, c->gram_part, c->max_part,
                               c->small.gram, c->num_sms, c->stream, &c->ls, c->batch());
    if (c->fast_boundary)
        return launch_gram_fast(src, c->Ecur(), c->T, c->d, c->m, inv_eps, c->n, c->gram_part, c->max_part,
                                c->small.gram, c->num_sms, c->stream, &c->ls);
    return launch_gram(src, c->Ecur(), c->T, c->d, c->m, inv_eps, c->n, c->gram_part, c->max_part, c->small.gram,
                       c->num_sms, c->stream, &c->ls, c->batch());
}

static cudaError_t do_apply(fds_ctx* c, int src, double inv_eps, double* Tout, double* E_out, double eps) {
    if (c->mma_boundary)
        return launch_apply_mma(src, c->Ecur(), c->T, c->d, c->small.A, c->m, inv_eps, c->n, Tout, E_out, c->P(0),
                                eps, c->num_sms, c->stream, &c->ls, c->batch());
    if (c->fast_boundary)
        return launch_apply_fast(src, c->Ecur(), c->T, c->d, c->A_host, c->m, inv_eps, c->n, Tout, E_out, c->P(0),
                                 eps, c->num_sms, c->stream, &c->ls);
    return launch_apply(src, c->Ecur(), c->T, c->d, c->small.A, c->m, inv_eps, c->n, Tout != nullptr ? Tout : c->T,
                        E_out, c->P(0), eps, c->stream, &c->ls, c->batch());
}


// ----------------------- in-library communicator (NCCL) ---------------------- //
#define FDS_NCCL(ctx, call)                                                                        \
    do {                                                                                           \
        int r__ = (ctx)->nccl->call;                                                               \
        if (r__ != 0) {                                                                            \
            (ctx)->err = std::string("nccl" #call ": ") + (ctx)->nccl->GetErrorString(r__);        \
            return 5;                                                                              \
        }                                                                                          \
    } while (0)

// ring halo exchange on the context's stream: the last 8h sites go to the right neighbour, the first 4h
// to the left one (tests/solvers/mock_fun3d/fun3d:29-37 does the same per RHS evaluation over MPI).
// With two ranks both neighbours are the same peer: sends and receives pair up in issue order.
static int halo_exchange(fds_ctx* c, int which, long long hs) {
    int64_t off[4], cnt[4];
    if (fds_halo_layout(c, which, hs, off, cnt) != 0) return fail(c, "halo_steps out of range");
    double* base = which == FDS_BUF_ENSEMBLE ? c->Ecur() : c->P(0);
    const int world = c->cfg.world, right = (c->cfg.rank + 1) % world, left = (c->cfg.rank + world - 1) % world;
    FDS_NCCL(c, GroupStart());
    FDS_NCCL(c, Send(base + off[0], (size_t)cnt[0], kNcclFloat64, right, c->comm, c->stream));
    FDS_NCCL(c, Send(base + off[1], (size_t)cnt[1], kNcclFloat64, left, c->comm, c->stream));
    FDS_NCCL(c, Recv(base + off[2], (size_t)cnt[2], kNcclFloat64, left, c->comm, c->stream));
    FDS_NCCL(c, Recv(base + off[3], (size_t)cnt[3], kNcclFloat64, right, c->comm, c->stream));
    FDS_NCCL(c, GroupEnd());
    c->ls.collectives++;
    if (which == FDS_BUF_ENSEMBLE) c->ens_halo_valid = hs; else c->pri_halo_valid = hs;
    return 0;
}

// sum of an (m+2)^2 Gram over the ranks (pascal_lite/operators/plinalg.py:19-44 reduces over MPI)
static int allreduce_gram(fds_ctx* c, double* g) {
    if (!c->comm) return 0;
    FDS_NCCL(c, AllReduce(g, g, (size_t)c->MX * c->MX, kNcclFloat64, kNcclSum, c->comm, c->stream));
    c->ls.collectives++;
    return 0;
}

// objective partials of the ranks -> global sums in c->jsum on every rank: all-gather, then the canonical
// tree continued over the ranks in a FIXED order (J does not depend on the number of GPUs)
static int reduce_jsum_ranks(fds_ctx* c, size_t len) {
    if (!c->comm || len == 0) return 0;
    const size_t need = len * (size_t)c->cfg.world;
    if (c->jgather_cap < need) {
        cudaFree(c->jgather);
        c->jgather = nullptr; c->jgather_cap = 0;
        FDS_CK(c, cudaMalloc(&c->jgather, need * sizeof(double)));
        c->jgather_cap = need;
    }
    FDS_NCCL(c, AllGather(c->jsum, c->jgather, len, kNcclFloat64, c->comm, c->stream));
    c->ls.collectives++;
    FDS_CK(c, launch_rank_tree_sum(c->jgather, c->cfg.world, (long long)len, c->jsum, c->stream, &c->ls));
    return 0;
}

extern "C" {

const char* fds_version(void) { return "fds_b200 0.1 (sm_100a)"; }

const char* fds_last_error(const fds_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int fds_ctx_create(const fds_config* cfg, fds_ctx** out) {
    *out = nullptr;
    if (!cfg || cfg->m < 1 || cfg->world < 1) { g_create_err = "bad config"; return 2; }
    const bool host_sys = cfg->system == FDS_SYS_HOST;
    const int tdim = host_sys ? 1 : toy_dim(cfg->system);
    if (host_sys && (cfg->n_local < 1 || cfg->world != 1 || cfg->batch > 1)) {
        g_create_err = "host-solver system: n_local >= 1, world = 1, no batch";
        return 2;
    }
    if (!host_sys && (tdim > 0 ? (cfg->n_local != tdim || cfg->world != 1) : cfg->n_local < 4)) {
        g_create_err = tdim > 0 ? "small ODE systems: n_local must equal the state dimension, world = 1"
                                : "bad config: n_local < 4";
        return 2;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                       " -- fds_b200 has no CPU fallback";
        return 3;
    }
    fds_ctx* c = new fds_ctx();
    c->cfg = *cfg;
    c->toy = tdim > 0;
    c->host_sys = host_sys;
    if (c->cfg.dt == 0.0) c->cfg.dt = 0.01;
    if (c->cfg.n_global == 0) c->cfg.n_global = cfg->n_local;
    c->m = cfg->m; c->M = cfg->m + 2; c->MO = cfg->m + 1; c->MX = cfg->m + 2;
    c->n = cfg->n_local;
#define CK0(call)                                                                \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            g_create_err = std::string(#call) + ": " + cudaGetErrorString(e__);  \
            fds_ctx_destroy(c);                                                  \
            return 1;                                                            \
        }                                                                        \
    } while (0)
    DeviceGuard guard__(cfg->device);
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != cfg->device) CK0(cudaSetDevice(cfg->device));   // reports a bad ordinal
    }
    CK0(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
    if (cfg->stream || (cfg->flags & FDS_FLAG_EXTERNAL_STREAM)) c->stream = (cudaStream_t)cfg->stream;
    else { CK0(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }

    long long H = cfg->max_halo_steps > 0 ? cfg->max_halo_steps : std::min<long long>(128, std::max<long long>(1, c->n / 64));
    if (H < 3) H = 3;                                   // the time-dilation look-ahead needs 3 steps
    if (cfg->world > 1 && 8 * H > c->n) {
        g_create_err = "n_local too small for the halo width (need 8*halo_steps <= n_local)";
        fds_ctx_destroy(c);
        return 2;
    }
    c->H = H;
    c->nb = cfg->batch > 1 ? cfg->batch : 1;
    c->nall = c->n;
    if (c->nb > 1) {
        // Batch: problems stacked along the site axis.  Between two problems lies a gap that holds the
        // right ghost cells of the first (4H sites) and the left ghost cells of the second (8H), so the
        // streaming kernel sweeps the whole stack as ONE array; a batched wrap refreshes the gaps.
        if (const char* e = getenv("FDS_BATCH_CTAS_PER_SM")) c->ctas_per_sm = atoi(e) == 1 ? 1 : 2;
        {   // two RK4 steps per pass (one 256-thread CTA per SM), as for a single problem
            const char* e = getenv("FDS_FUSE2");
            c->fuse2 = (e ? e[0] != '0' : true) && (cfg->system == FDS_SYS_L96_RK4 || cfg->system == FDS_SYS_L96_EULER) &&
                       two_step_runs_per_cta(c->M) >= 1;
            if (c->fuse2) c->ctas_per_sm = 1;
            c->fuse4 = c->fuse2 && cfg->system == FDS_SYS_L96_EULER && H >= 4 && !(e && e[0] == '2');
        }
        // strips per CTA: two CTAs of <= 256 threads per SM, or one of <= 480 with a producer warp
        const int R2 = c->fuse2 ? two_step_runs_per_cta(c->M) : (c->ctas_per_sm == 2 ? 256 / c->M : stream_runs_per_cta(c->M));
        if ((cfg->system != FDS_SYS_L96_RK4 && cfg->system != FDS_SYS_L96_EULER) || cfg->world != 1 ||
            c->n % 64 != 0 || R2 < 1 ||
            (cfg->flags & FDS_FLAG_SIMPLE_KERNELS)) {
            g_create_err = "batch: needs a Lorenz-96 system, world = 1, n_local a multiple of 64, m <= 126";
            fds_ctx_destroy(c);
            return 2;
        }
        H = cfg->max_halo_steps > 0 ? cfg->max_halo_steps : 16;
        H = std::max<long long>(8, (H + 7) / 8 * 8);       // 8H a multiple of the 64-site leaf
        if (8 * H > c->n) H = std::max<long long>(8, c->n / 64 * 8);
        c->H = H;
        c->gl0 = 8 * H;
        const long long slots_per_wave = (long long)c->ctas_per_sm * c->num_sms * R2;   // strips resident at once
        c->spp = (int)std::max<long long>(1, std::min<long long>(8, (slots_per_wave + c->nb / 2) / c->nb));
        while (c->spp > 1 && (c->n + 12 * H) / c->spp < 256) --c->spp;        // keep strips long
        const long long unit = 64LL * c->spp;
        c->slot = (c->n + 12 * H + unit - 1) / unit * unit;
        c->nall = c->slot * c->nb;
        const long long nstrips = (long long)c->nb * c->spp;
        const long long waves = (nstrips + slots_per_wave - 1) / slots_per_wave;
        c->runs_per_cta = (int)std::min<long long>(R2, (nstrips + waves * c->ctas_per_sm * c->num_sms - 1) / (waves * c->ctas_per_sm * c->num_sms));
        c->stream_ok = true;
        StripPlan& pl = c->plan;
        pl.a0 = 0; pl.ls = c->slot / c->spp; pl.nstrips = nstrips; pl.nblk = (int)(pl.ls / 8);
        pl.alo = -kPlanPad; pl.ahi = c->nall + kPlanPad; pl.node_shift = kLeafShift; pl.nnode = c->nall / kLeafSites;
        c->alo = pl.alo; c->ahi = pl.ahi;
        c->ds.assign(c->nb, 0.0);
        c->Fhost.assign(c->nb, make_double2(8.0, 8.0));
    } else {
        int R = stream_runs_per_cta(c->M);
        c->stream_ok = R >= 1 && (cfg->system == FDS_SYS_L96_RK4 || cfg->system == FDS_SYS_L96_EULER) &&
                       !(cfg->flags & FDS_FLAG_SIMPLE_KERNELS);
        // temporal blocking: two RK4 steps per pass halve the HBM traffic per step; the kernel then holds
        // twice the pipeline state per thread, so a CTA carries 256 / M strips instead of 480 / M
        {
            const char* e = getenv("FDS_FUSE2");
            const bool want = e ? e[0] != '0' : true;
            c->fuse2 = want && c->stream_ok && two_step_runs_per_cta(c->M) >= 1 && H >= 2;   // RK4 and forward Euler
            if (c->fuse2) { R = two_step_runs_per_cta(c->M); c->runs_per_cta = R; }
            c->fuse4 = c->fuse2 && cfg->system == FDS_SYS_L96_EULER && H >= 4 && !(e && e[0] == '2');   // FDS_FUSE2=2: at most two
        }
        long long max_strips = (long long)c->num_sms * (R >= 1 ? R : 1);
        // the two-step kernel packs (strip, column) pairs over all 256 lanes of a CTA (l96_step.cu)
        if (c->fuse2 && c->M >= 9) max_strips = (long long)c->num_sms * 256 / c->M;
        c->plan = make_strip_plan(c->n, H, max_strips);
        c->alo = c->plan.alo;
        c->ahi = c->plan.ahi;
    }
    const size_t sites = (size_t)(c->ahi - c->alo);
    for (int j = 0; j < 2; ++j) {
        CK0(cudaMalloc(&c->Ebase[j], sites * c->M * sizeof(double)));
        CK0(cudaMemsetAsync(c->Ebase[j], 0, sites * c->M * sizeof(double), c->stream));
    }
    for (int j = 0; j < 4; ++j) {
        CK0(cudaMalloc(&c->Pbase[j], sites * sizeof(double)));
        CK0(cudaMemsetAsync(c->Pbase[j], 0, sites * sizeof(double), c->stream));
    }
    // +2 doubles: the boundary kernels fetch tiles with 16-byte-granular bulk copies
    CK0(cudaMalloc(&c->T_alloc, ((size_t)c->nall * c->MO + 2) * sizeof(double)));
    CK0(cudaMemsetAsync(c->T_alloc, 0, (size_t)c->nall * c->MO * sizeof(double), c->stream));
    CK0(cudaMalloc(&c->d_alloc, ((size_t)c->nall + 2) * sizeof(double)));
    CK0(cudaMemsetAsync(c->d_alloc, 0, (size_t)c->nall * sizeof(double), c->stream));
    c->T = c->T_alloc + c->gl0 * c->MO;
    c->d = c->d_alloc + c->gl0;
    c->max_part = c->nb > 1 ? 4 : 2 * c->num_sms;
    CK0(cudaMalloc(&c->gram_part, (size_t)c->nb * c->max_part * c->MX * c->MX * sizeof(double)));
    if (c->nb > 1) CK0(cudaMalloc(&c->Fpp, (size_t)c->nb * sizeof(double2)));
    {   // small matrices in one block (per problem)
        const size_t mm = (size_t)c->m * c->m;
        const size_t tot = (2 * (size_t)c->MX * c->MX + (size_t)c->MO * c->MX + 2 * mm + c->m + c->MO + 4 + 1) & ~(size_t)1;
        c->small_stride = (long long)tot;
        CK0(cudaMalloc(&c->small_block, tot * c->nb * sizeof(double)));
        CK0(cudaMemsetAsync(c->small_block, 0, tot * c->nb * sizeof(double), c->stream));
        double* p = c->small_block;
        c->small.gram = p; p += (size_t)c->MX * c->MX;
        c->small.A = p;    p += (size_t)c->MO * c->MX;
        c->small.L1 = p;   p += mm;
        c->small.R = p;    p += mm;
        c->small.b = p;    p += c->m;
        c->small.gdil = p; p += c->MO;
        c->small.status = reinterpret_cast<int*>(p); p += 4;      // [int status, int two_pass][double cond][2 spare]
        c->gram2 = p;                                             // second-pass Gram of the gated CholeskyQR2
    }
    c->bd.count = c->nb; c->bd.slot = c->slot; c->bd.small_stride = c->small_stride;
    c->bd.part_stride = (long long)c->max_part * c->MX * c->MX;
    c->fast_boundary = apply_fast_supported(c->m) && !(cfg->flags & FDS_FLAG_SIMPLE_KERNELS) && c->nb == 1;
    c->mma_boundary = gram_mma_supported(c->m) && apply_mma_supported(c->m) && !(cfg->flags & FDS_FLAG_SIMPLE_KERNELS) &&
                      !c->toy;
    CK0(cudaMallocHost(&c->A_host, (size_t)c->MO * c->MX * sizeof(double)));
    CK0(cudaMallocHost(&c->out_host, (size_t)c->nb * ((size_t)c->m * c->m + 2 * c->m + 3) * sizeof(double)));
    // the boundary runs without a host round trip when both big passes are tensor-core kernels (m = 16, 32)
    {
        const char* e = getenv("FDS_GATED");
        c->gated = c->mma_boundary && !(e && e[0] == '0');
    }
    CK0(cudaMalloc(&c->c4, 4 * sizeof(double)));
    {
        const double w[4] = {-11.0 / 6.0, 3.0, -1.5, 1.0 / 3.0};
        CK0(cudaMemcpyAsync(c->c4, w, sizeof(w), cudaMemcpyHostToDevice, c->stream));
        CK0(cudaStreamSynchronize(c->stream));
    }
    c->nleaf = (c->nall + 63) / 64;
    CK0(cudaMalloc(&c->jpart, (size_t)c->nleaf * c->M * 2 * sizeof(double)));
    CK0(cudaMalloc(&c->jscratch, (size_t)2 * ((c->nleaf + 255) / 256) * c->M * 2 * sizeof(double)));
    if (small_smem_bytes(c->m) > 220 * 1024 || lss_smem_bytes(c->m) > 220 * 1024) {
        g_create_err = "subspace dimension too large for the single-CTA kernels (m <= 80)";
        fds_ctx_destroy(c);
        return 2;
    }
    CK0(cudaStreamSynchronize(c->stream));
#undef CK0
    *out = c;
    return 0;
}

void fds_ctx_destroy(fds_ctx* c) {
    if (!c) return;
    FDS_ON_DEVICE(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->own_comm && c->nccl) c->nccl->CommDestroy(c->comm);
    cudaFree(c->jgather);
    if (c->out_host) cudaFreeHost(c->out_host);
    for (int j = 0; j < 2; ++j) cudaFree(c->Ebase[j]);
    for (int j = 0; j < 4; ++j) cudaFree(c->Pbase[j]);
    cudaFree(c->T_alloc); cudaFree(c->d_alloc); cudaFree(c->Fpp); cudaFree(c->sens); cudaFree(c->gram_part); cudaFree(c->small_block); cudaFree(c->c4);
    cudaFree(c->jpart); cudaFree(c->jscratch); cudaFree(c->jsum);
    for (int j = 0; j < 2; ++j) {
        cudaFree(c->stage[j]);
        if (c->ev_copied[j]) cudaEventDestroy(c->ev_copied[j]);
        if (c->ev_done[j]) cudaEventDestroy(c->ev_done[j]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaFree(c->jnode); cudaFree(c->jnode_scratch); cudaFree(c->prof_cycles);
    cudaFreeHost(c->A_host);
    for (auto& pe : c->prof_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// pinned host memory for callers that want full-speed PCIe transfers (any host buffer works)
void* fds_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void fds_host_free(void* p) { if (p) cudaFreeHost(p); }

int fds_sync(fds_ctx* c) {
    FDS_ON_DEVICE(c);
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------- communicator -------------------------------- //
int fds_comm_unique_id(void* id128) {
    std::string err;
    const NcclApi* api = nccl_api(&err);
    if (!api) { g_create_err = err; return 5; }
    const int r = api->GetUniqueId(reinterpret_cast<NcclUniqueId*>(id128));
    if (r != 0) { g_create_err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return 5; }
    return 0;
}

int fds_comm_init(fds_ctx* c, const void* id128) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world < 2) return fail(c, "fds_comm_init: the context was created with world = 1");
    if (c->comm) return fail(c, "fds_comm_init: a communicator is already attached");
    c->nccl = nccl_api(&c->err);
    if (!c->nccl) return 5;
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    FDS_NCCL(c, CommInitRank(&c->comm, c->cfg.world, id, c->cfg.rank));
    c->own_comm = true;
    return 0;
}

int fds_comm_attach(fds_ctx* c, void* nccl_comm) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world < 2) return fail(c, "fds_comm_attach: the context was created with world = 1");
    if (c->comm) return fail(c, "fds_comm_attach: a communicator is already attached");
    c->nccl = nccl_api(&c->err);
    if (!c->nccl) return 5;
    int count = 0, rank = -1;
    FDS_NCCL(c, CommCount(nccl_comm, &count));
    FDS_NCCL(c, CommUserRank(nccl_comm, &rank));
    if (count != c->cfg.world || rank != c->cfg.rank) return fail(c, "fds_comm_attach: communicator size / rank differ from the context's");
    c->comm = nccl_comm;
    c->own_comm = false;
    return 0;
}

int fds_comm_attached(const fds_ctx* c) { return c->comm != nullptr; }
int64_t fds_collectives(const fds_ctx* c) { return c->ls.collectives; }

// ------------------------------- introspection ------------------------------- //
int fds_get_ensemble_rows(fds_ctx* c, int64_t site0, int64_t nsites, double* out) {
    FDS_ON_DEVICE(c);
    if (c->nb > 1) return fail(c, "fds_get_ensemble_rows: not available for a batch");
    if (nsites < 0 || site0 < c->alo || site0 + nsites > c->ahi) return fail(c, "fds_get_ensemble_rows: range outside the allocated sites");
    FDS_CK(c, cudaMemcpyAsync(out, c->Ecur() + site0 * c->M, (size_t)nsites * c->M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_get_ensemble_column(fds_ctx* c, int k, double* out) {
    FDS_ON_DEVICE(c);
    if (c->nb > 1) return fail(c, "fds_get_ensemble_column: not available for a batch");
    if (k < 0 || k >= c->M) return fail(c, "fds_get_ensemble_column: column out of range");
    // P(3) is scratch outside fds_boundary_dxdt (the third look-ahead state of the time-dilation stencil)
    FDS_CK(c, launch_extract_col0(c->Ecur() + k, c->M, c->P(3), c->n, c->stream, &c->ls));
    FDS_CK(c, cudaMemcpyAsync(out, c->P(3), (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_strip_plan(const fds_ctx* c, int64_t out[8]) {
    if (!c->stream_ok) return 2;
    out[0] = c->plan.a0; out[1] = c->plan.ls; out[2] = c->plan.nstrips; out[3] = c->plan.nblk;
    out[4] = c->plan.node_shift; out[5] = c->plan.nnode;
    out[6] = (c->fuse2 && c->M >= 9 && c->nb == 1) ? 256 : 0;    // lanes a CTA packs (strip, column) pairs over; 0 = whole strips per CTA
    out[7] = c->fuse4 ? 4 : (c->fuse2 ? 2 : 1);
    return 0;
}

int fds_fp64_issue_rate(int device, double* winst_per_s, double* seconds, double* winst_per_clk_sm) {
    int ndev = 0, sms = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) { g_create_err = "fds_fp64_issue_rate: no such CUDA device"; return 3; }
    DeviceGuard guard__(device);
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = fp64_issue_rate(sms, nullptr, winst_per_s, seconds, winst_per_clk_sm);
    if (e != cudaSuccess) { g_create_err = std::string("fds_fp64_issue_rate: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

// write-combined page-locked memory: for buffers the host only WRITES and the device reads (uploads)
void* fds_host_alloc_wc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

int64_t fds_kernel_launches(const fds_ctx* c) { return c->ls.launches; }

int64_t fds_profile_steps(const fds_ctx* c) { return c->prof_steps; }

int fds_profile_enable(fds_ctx* c, int enabled) {
    FDS_ON_DEVICE(c);
    if (enabled) {
        if (!c->prof_cycles) FDS_CK(c, cudaMalloc(&c->prof_cycles, 2 * sizeof(unsigned long long)));
        FDS_CK(c, cudaMemsetAsync(c->prof_cycles, 0, 2 * sizeof(unsigned long long), c->stream));
    }
    c->profiling = enabled != 0;
    c->prof_used = 0;
    c->prof_steps = 0;
    return 0;
}

int fds_profile_read(fds_ctx* c, double* total_ms, int64_t* launches) {
    FDS_ON_DEVICE(c);
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    double tot = 0.0;
    for (size_t j = 0; j < c->prof_used; ++j) {
        float ms = 0.f;
        FDS_CK(c, cudaEventElapsedTime(&ms, c->prof_events[j].first, c->prof_events[j].second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = (int64_t)c->prof_used;
    c->prof_used = 0;
    return 0;
}

// mean SM cycles a CTA of the profiled step-kernel launches ran (clock64 inside the kernel): with the FP64
// instruction count this gives the pipe utilisation at the kernel's OWN clock, whatever the power cap did to it
int fds_profile_cycles(fds_ctx* c, double* mean_cycles_per_cta) {
    FDS_ON_DEVICE(c);
    unsigned long long h[2] = {0, 0};
    if (c->prof_cycles) {
        FDS_CK(c, cudaMemcpyAsync(h, c->prof_cycles, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        FDS_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (mean_cycles_per_cta) *mean_cycles_per_cta = h[1] ? (double)h[0] / (double)h[1] : 0.0;
    return 0;
}

void* fds_buffer_ptr(const fds_ctx* c, int which) {
    switch (which) {
        case FDS_BUF_ENSEMBLE: return c->Ecur();
        case FDS_BUF_PRIMAL: return c->P(0);
        case FDS_BUF_GRAM: return c->small.gram;
        case FDS_BUF_JSUM: return c->jsum;
        case FDS_BUF_TANGENT: return c->T;
        case FDS_BUF_DXDT: return c->d;
    }
    return nullptr;
}

int fds_set_dxdt_weights(fds_ctx* c, const double w[4]) {
    FDS_ON_DEVICE(c);
    FDS_CK(c, cudaMemcpyAsync(c->c4, w, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    c->stencil.order = 3;
    for (int j = 0; j < 4; ++j) c->stencil.c[j] = w[j];
    return 0;
}

// TimeDilation.order_of_accuracy (fds/timedilation.py:8-12,64): `order` look-ahead steps, order + 1 weights
int fds_set_dxdt_stencil(fds_ctx* c, int order, const double* w) {
    if (order == 3) return fds_set_dxdt_weights(c, w);
    if (order < 1 || order > kMaxDilationOrder) return fail(c, "fds_set_dxdt_stencil: order must be 1 .. 8");
    if (c->toy || c->host_sys) return fail(c, "fds_set_dxdt_stencil: small ODE systems and host solvers implement order 3 only");
    if (order > c->H) return fail(c, "fds_set_dxdt_stencil: order exceeds the halo capacity (max_halo_steps)");
    c->stencil.order = order;
    for (int j = 0; j <= order; ++j) c->stencil.c[j] = w[j];
    return 0;
}

int fds_dilation_order(const fds_ctx* c) { return c->stencil.order; }

int fds_set_time_dilation(fds_ctx* c, int mode) {
    if (mode < 0 || mode > 3) return fail(c, "fds_set_time_dilation: mode must be 0 .. 3");
    if (mode == FDS_DILATION_ANALYTIC && (c->toy || c->host_sys))
        return fail(c, "fds_set_time_dilation: the analytic derivative exists for the Lorenz-96 systems only");
    c->dil_mode = mode;
    c->time_dilation = mode != 0;
    return 0;
}

int fds_set_dxdt(fds_ctx* c, const double* d_host) {
    FDS_ON_DEVICE(c);
    FDS_CK(c, cudaMemcpy2DAsync(c->d, (size_t)(c->nb > 1 ? c->slot : c->n) * sizeof(double), d_host,
                                (size_t)c->n * sizeof(double), (size_t)c->n * sizeof(double), c->nb,
                                cudaMemcpyHostToDevice, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_set_qr_passes(fds_ctx* c, int passes) {
    if (passes != 0 && passes != 2) return fail(c, "fds_set_qr_passes: 0 (adaptive) or 2 (always CholeskyQR2)");
    c->qr_mode = passes;
    return 0;
}

double fds_last_condition(const fds_ctx* c) { return c->cond_last; }

int fds_single_pass_pending(const fds_ctx* c) { return c->qr_single ? 1 : 0; }

int fds_last_qr_passes(const fds_ctx* c) { return c->passes_last; }

// ------------------------------- state in / out ------------------------------ //
static int sync_primal_from_ensemble(fds_ctx* c) {
    if (c->state_in_ensemble) {
        if (!c->primal_synced)
            FDS_CK(c, launch_extract_col0(c->E_all(c->cur), c->M, c->P_all(0), c->nall, c->stream, &c->ls));
        c->state_in_ensemble = false;
        c->pri_halo_valid = 0;
    }
    return 0;
}

int fds_set_state(fds_ctx* c, const double* u) {
    FDS_ON_DEVICE(c);
    // batch: u is [nb][n]; rows land `slot` sites apart
    FDS_CK(c, cudaMemcpy2DAsync(c->P(0), (size_t)(c->nb > 1 ? c->slot : c->n) * sizeof(double), u,
                                (size_t)c->n * sizeof(double), (size_t)c->n * sizeof(double), c->nb,
                                cudaMemcpyHostToDevice, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    c->state_in_ensemble = false;
    c->primal_synced = false;
    c->pri_halo_valid = 0;
    return 0;
}

int fds_get_state(fds_ctx* c, double* u) {
    FDS_ON_DEVICE(c);
    if (int r = sync_primal_from_ensemble(c)) return r;
    FDS_CK(c, cudaMemcpy2DAsync(u, (size_t)c->n * sizeof(double), c->P(0),
                                (size_t)(c->nb > 1 ? c->slot : c->n) * sizeof(double), (size_t)c->n * sizeof(double),
                                c->nb, cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

static int ensure_stage(fds_ctx* c) {
    if (c->stage[0]) return 0;
    c->stage_sites = std::min<long long>(c->n, 1LL << 18);
    FDS_CK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j) {
        FDS_CK(c, cudaMalloc(&c->stage[j], (size_t)c->stage_sites * c->M * sizeof(double)));
        FDS_CK(c, cudaEventCreateWithFlags(&c->ev_copied[j], cudaEventDisableTiming));
        FDS_CK(c, cudaEventCreateWithFlags(&c->ev_done[j], cudaEventDisableTiming));
    }
    return 0;
}

static int materialize_tangents(fds_ctx* c) {
    if (c->tangents_implicit) {
        FDS_CK(c, launch_diff_tangents(c->E_all(c->cur), c->M, c->eps_last, c->T_alloc, c->nall, c->stream, &c->ls));
        c->tangents_implicit = false;
    }
    return 0;
}

// Host rows V[j][.] (j < m), v[.] -> device T[site][m+1], in chunks of stage_sites sites: the
// copy stream uploads chunk i+1 while the context's stream transposes chunk i.
// batch: V is [nb][m][n], v is [nb][n]
int fds_set_tangents(fds_ctx* c, const double* Vall, const double* vall) {
    FDS_ON_DEVICE(c);
    if (int r = ensure_stage(c)) return r;
    FDS_CK(c, cudaStreamSynchronize(c->stream));       // T and the stages are free
    long long i = 0;
    for (int pb = 0; pb < c->nb; ++pb) {
    const double* V = Vall + (size_t)pb * c->m * c->n;
    const double* v = vall ? vall + (size_t)pb * c->n : nullptr;
    double* Tp = c->T + (size_t)pb * c->slot * c->MO;
    for (long long s0 = 0; s0 < c->n; s0 += c->stage_sites, ++i) {
        const int b = (int)(i & 1);
        const long long cnt = std::min(c->stage_sites, c->n - s0);
        double* stg = c->stage[b];
        if (i >= 2) FDS_CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        // stage[j][0..cnt) <- V[j][s0..s0+cnt), row m <- v
        FDS_CK(c, cudaMemcpy2DAsync(stg, (size_t)c->stage_sites * sizeof(double), V + s0,
                                    (size_t)c->n * sizeof(double), (size_t)cnt * sizeof(double), c->m,
                                    cudaMemcpyHostToDevice, c->copy_stream));
        if (v) FDS_CK(c, cudaMemcpyAsync(stg + (size_t)c->m * c->stage_sites, v + s0, (size_t)cnt * sizeof(double),
                                         cudaMemcpyHostToDevice, c->copy_stream));
        else FDS_CK(c, cudaMemsetAsync(stg + (size_t)c->m * c->stage_sites, 0, (size_t)cnt * sizeof(double), c->copy_stream));
        FDS_CK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
        FDS_CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        FDS_CK(c, launch_transpose_back(stg, c->stage_sites, cnt, c->MO, Tp + (size_t)s0 * c->MO, c->stream, &c->ls));
        FDS_CK(c, cudaEventRecord(c->ev_done[b], c->stream));
    }
    }
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    c->tangents_implicit = false;
    return 0;
}

int fds_get_tangents(fds_ctx* c, double* Vall, double* vall) {
    FDS_ON_DEVICE(c);
    if (int r = materialize_tangents(c)) return r;
    if (int r = ensure_stage(c)) return r;
    long long i = 0;
    for (int pb = 0; pb < c->nb; ++pb) {
    double* V = Vall ? Vall + (size_t)pb * c->m * c->n : nullptr;
    double* v = vall ? vall + (size_t)pb * c->n : nullptr;
    const double* Tp = c->T + (size_t)pb * c->slot * c->MO;
    for (long long s0 = 0; s0 < c->n; s0 += c->stage_sites, ++i) {
        const int b = (int)(i & 1);
        const long long cnt = std::min(c->stage_sites, c->n - s0);
        double* stg = c->stage[b];
        if (i >= 2) FDS_CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));   // download of chunk i-2 done
        FDS_CK(c, launch_transpose(Tp + (size_t)s0 * c->MO, cnt, c->MO, stg, c->stage_sites, c->stream, &c->ls));
        FDS_CK(c, cudaEventRecord(c->ev_done[b], c->stream));
        FDS_CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        if (V) FDS_CK(c, cudaMemcpy2DAsync(V + s0, (size_t)c->n * sizeof(double), stg,
                                           (size_t)c->stage_sites * sizeof(double), (size_t)cnt * sizeof(double), c->m,
                                           cudaMemcpyDeviceToHost, c->copy_stream));
        if (v) FDS_CK(c, cudaMemcpyAsync(v + s0, stg + (size_t)c->m * c->stage_sites, (size_t)cnt * sizeof(double),
                                         cudaMemcpyDeviceToHost, c->copy_stream));
        FDS_CK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    }
    }
    FDS_CK(c, cudaStreamSynchronize(c->copy_stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- host-solver mode: the ensemble itself crosses the boundary --------------------------- //
// E_host[M][n] (row k = trajectory k) <-> device [n][M], same staged transposes as the tangents
int fds_get_ensemble(fds_ctx* c, double* Eh) {
    FDS_ON_DEVICE(c);
    if (c->nb > 1) return fail(c, "fds_get_ensemble: not available for a batch");
    if (int r = ensure_stage(c)) return r;
    long long i = 0;
    for (long long s0 = 0; s0 < c->n; s0 += c->stage_sites, ++i) {
        const int b = (int)(i & 1);
        const long long cnt = std::min(c->stage_sites, c->n - s0);
        double* stg = c->stage[b];
        if (i >= 2) FDS_CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        FDS_CK(c, launch_transpose(c->Ecur() + (size_t)s0 * c->M, cnt, c->M, stg, c->stage_sites, c->stream, &c->ls));
        FDS_CK(c, cudaEventRecord(c->ev_done[b], c->stream));
        FDS_CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        FDS_CK(c, cudaMemcpy2DAsync(Eh + s0, (size_t)c->n * sizeof(double), stg, (size_t)c->stage_sites * sizeof(double),
                                    (size_t)cnt * sizeof(double), c->M, cudaMemcpyDeviceToHost, c->copy_stream));
        FDS_CK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    }
    FDS_CK(c, cudaStreamSynchronize(c->copy_stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// the advanced ensemble comes back: column 0 is the new primal state, the tangents are implicit
// in it, (E_k - E_0) / eps, exactly as after fds_segment_advance
int fds_set_ensemble(fds_ctx* c, const double* Eh, double eps) {
    FDS_ON_DEVICE(c);
    if (c->nb > 1) return fail(c, "fds_set_ensemble: not available for a batch");
    if (int r = ensure_stage(c)) return r;
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    long long i = 0;
    for (long long s0 = 0; s0 < c->n; s0 += c->stage_sites, ++i) {
        const int b = (int)(i & 1);
        const long long cnt = std::min(c->stage_sites, c->n - s0);
        double* stg = c->stage[b];
        if (i >= 2) FDS_CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        FDS_CK(c, cudaMemcpy2DAsync(stg, (size_t)c->stage_sites * sizeof(double), Eh + s0, (size_t)c->n * sizeof(double),
                                    (size_t)cnt * sizeof(double), c->M, cudaMemcpyHostToDevice, c->copy_stream));
        FDS_CK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
        FDS_CK(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        FDS_CK(c, launch_transpose_back(stg, c->stage_sites, cnt, c->M, c->Ecur() + (size_t)s0 * c->M, c->stream, &c->ls));
        FDS_CK(c, cudaEventRecord(c->ev_done[b], c->stream));
    }
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    c->state_in_ensemble = true;
    c->primal_synced = false;
    c->tangents_implicit = true;
    c->eps_last = eps;
    c->ens_halo_valid = 0;
    return 0;
}

// time dilation of the host-solver mode: the states after 1, 2, 3 steps (fds/timedilation.py:63-82)
int fds_set_primal_history(fds_ctx* c, int j, const double* u_host) {
    FDS_ON_DEVICE(c);
    if (j < 1 || j > 3) return fail(c, "fds_set_primal_history: j must be 1, 2 or 3");
    FDS_CK(c, cudaMemcpyAsync(c->P(j), u_host, (size_t)c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_random_tangents(fds_ctx* c, uint64_t seed) {
    FDS_ON_DEVICE(c);
    if (c->nb > 1) FDS_CK(c, launch_fill_random(c->T_alloc, c->nall, c->MO, c->m, 0, c->nall, seed, c->stream, &c->ls));
    else FDS_CK(c, launch_fill_random(c->T, c->n, c->MO, c->m, c->cfg.site_offset, c->cfg.n_global, seed, c->stream, &c->ls));
    c->tangents_implicit = false;
    return 0;
}

// ----------------------------------- halos ----------------------------------- //
int64_t fds_halo_steps(const fds_ctx* c, int64_t steps) { return c->toy ? steps : std::min<long long>(steps, c->H); }

int fds_halo_wrap_ensemble(fds_ctx* c, int64_t hs) {
    FDS_ON_DEVICE(c);
    if (c->toy) { c->ens_halo_valid = hs; return 0; }
    if (c->cfg.world != 1 && c->comm) return halo_exchange(c, FDS_BUF_ENSEMBLE, hs);
    if (c->cfg.world != 1) return fail(c, "fds_halo_wrap_ensemble: world > 1, attach a communicator (fds_comm_init) or exchange halos between ranks yourself");
    if (hs < 1 || hs > c->H) return fail(c, "halo_steps out of range");
    if (c->nb > 1) FDS_CK(c, launch_halo_wrap_batch(c->Ecur(), c->M, c->n, 8 * hs, 4 * hs, c->nb, c->slot, c->stream, &c->ls));
    else
    FDS_CK(c, launch_halo_wrap(c->Ecur(), c->M, c->n, 8 * hs, 4 * hs, c->stream, &c->ls));
    c->ens_halo_valid = hs;
    return 0;
}

int fds_halo_wrap_primal(fds_ctx* c, int64_t hs) {
    FDS_ON_DEVICE(c);
    if (c->toy) { c->pri_halo_valid = hs; return 0; }
    if (c->cfg.world != 1 && c->comm) return halo_exchange(c, FDS_BUF_PRIMAL, hs);
    if (c->cfg.world != 1) return fail(c, "fds_halo_wrap_primal: world > 1, attach a communicator (fds_comm_init) or exchange halos between ranks yourself");
    if (hs < 1 || hs > c->H) return fail(c, "halo_steps out of range");
    if (c->nb > 1) FDS_CK(c, launch_halo_wrap_batch(c->P(0), 1, c->n, 8 * hs, 4 * hs, c->nb, c->slot, c->stream, &c->ls));
    else
    FDS_CK(c, launch_halo_wrap(c->P(0), 1, c->n, 8 * hs, 4 * hs, c->stream, &c->ls));
    c->pri_halo_valid = hs;
    return 0;
}

int fds_halo_mark_valid(fds_ctx* c, int which, int64_t hs) {
    if (hs < 0 || hs > c->H) return fail(c, "halo_steps out of range");
    if (which == FDS_BUF_ENSEMBLE) c->ens_halo_valid = hs;
    else if (which == FDS_BUF_PRIMAL) c->pri_halo_valid = hs;
    else return fail(c, "fds_halo_mark_valid: buffer has no halo");
    return 0;
}

int fds_halo_layout(const fds_ctx* c, int which, int64_t hs, int64_t off[4], int64_t cnt[4]) {
    const long long W = which == FDS_BUF_ENSEMBLE ? c->M : 1;
    if (which != FDS_BUF_ENSEMBLE && which != FDS_BUF_PRIMAL) return 2;
    if (hs < 1 || hs > c->H || 8 * hs > c->n) return 2;
    const long long gl = 8 * hs, gr = 4 * hs;
    off[0] = (c->n - gl) * W; cnt[0] = gl * W;   // send to right neighbour (its left ghost)
    off[1] = 0;               cnt[1] = gr * W;   // send to left neighbour (its right ghost)
    off[2] = -gl * W;         cnt[2] = gl * W;   // receive from left neighbour
    off[3] = c->n * W;        cnt[3] = gr * W;   // receive from right neighbour
    return 0;
}

// ---------------------------------- stepping --------------------------------- //
static int ensure_jsum(fds_ctx* c, long long steps) {
    if (steps <= c->jsum_cap) return 0;
    double* nj = nullptr;
    const long long cap = std::max<long long>(steps, 2 * c->jsum_cap);
    const size_t per_step = (size_t)c->nb * c->M * 2;          // [step][problem][column][2]
    FDS_CK(c, cudaMalloc(&nj, cap * per_step * sizeof(double)));
    FDS_CK(c, cudaMemsetAsync(nj, 0, cap * per_step * sizeof(double), c->stream));
    if (c->jsum) {
        FDS_CK(c, cudaMemcpyAsync(nj, c->jsum, (size_t)c->jsum_cap * per_step * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        FDS_CK(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->jsum);
    }
    c->jsum = nj;
    c->jsum_cap = cap;
    if (c->stream_ok && !c->jnode) {
        // node partials of one halo chunk (<= H steps) at a time, reduced right after it: the buffer does
        // not grow with the number of steps of a segment
        const size_t row = (size_t)c->plan.nnode * c->M * 2;
        FDS_CK(c, cudaMalloc(&c->jnode, (size_t)c->H * row * sizeof(double)));
        if (c->nb == 1)
            FDS_CK(c, cudaMalloc(&c->jnode_scratch,
                                 (size_t)2 * c->H * ((c->plan.nnode + 255) / 256) * c->M * 2 * sizeof(double)));
    }
    return 0;
}

// one step of `M_` columns X -> Y with `valid` halo steps left; objective sums of step `jstep`
// (streaming kernel: node partials into jnode[jstep], reduced later in one batch by
// reduce_nodes(); simple kernels: reduced right away into jsum[jstep])
static int one_step(fds_ctx* c, const double* X, double* Y, int M_, double F0, double Flast, long long valid,
                    long long jstep, long long jslot, int nfuse = 1, double* primal_out = nullptr) {
    const bool two = nfuse >= 2;
    double* jsum_slot = c->jsum + (size_t)jstep * M_ * 2;
    StepArgs a;
    a.X = X; a.Y = Y; a.M = M_; a.n = c->nall;
    if (c->nb > 1) { a.Fpp = c->Fpp; a.spp = c->spp; a.slot = c->slot; a.runs_per_cta = c->runs_per_cta; a.ctas_per_sm = c->ctas_per_sm; }
    else if (c->fuse2) a.runs_per_cta = c->runs_per_cta;
    a.lo = -8 * (valid - 1); a.hi = c->n + 4 * (valid - 1);
    a.F0 = F0; a.Flast = Flast; a.dt = c->cfg.dt; a.exact = c->cfg.math == FDS_MATH_EXACT;
    a.jpart = jsum_slot ? c->jpart : nullptr;
    a.nleaf = c->nleaf;
    const bool euler = c->cfg.system == FDS_SYS_L96_EULER;
    if (c->stream_ok && M_ == c->M) {
        a.jpart = c->jnode + (size_t)jslot * c->plan.nnode * c->M * 2;
        if (two) { a.two_steps = 1; a.jpart2 = a.jpart + (size_t)c->plan.nnode * c->M * 2; a.nfuse = nfuse; }
        a.primal_out = primal_out;
        a.cycles = c->profiling ? c->prof_cycles : nullptr;
        cudaEvent_t e1 = nullptr;
        if (c->profiling) {
            if (c->prof_used == c->prof_events.size()) {
                cudaEvent_t a0, a1;
                FDS_CK(c, cudaEventCreate(&a0));
                FDS_CK(c, cudaEventCreate(&a1));
                c->prof_events.emplace_back(a0, a1);
            }
            FDS_CK(c, cudaEventRecord(c->prof_events[c->prof_used].first, c->stream));
            e1 = c->prof_events[c->prof_used].second;
            c->prof_used++;
            c->prof_steps += nfuse;
        }
        FDS_CK(c, launch_l96_rk4_stream(a, c->plan, c->stream, &c->ls, euler ? 1 : 0));
        if (e1) FDS_CK(c, cudaEventRecord(e1, c->stream));
    } else {
        FDS_CK(c, launch_l96_simple(a, euler ? 1 : 0, c->stream, &c->ls));
        FDS_CK(c, launch_stats_leaf(Y, M_, c->n, c->jpart, c->nleaf, c->stream, &c->ls));
        FDS_CK(c, launch_tree_reduce(c->jpart, c->nleaf, 2 * M_, jsum_slot, c->jscratch, c->stream, &c->ls));
    }
    return 0;
}

// node partials of steps [step0, step0 + nsteps) -> jsum, one batched canonical-tree reduction
static int reduce_nodes(fds_ctx* c, long long step0, long long nsteps) {
    if (!c->stream_ok || nsteps < 1) return 0;
    if (c->nb > 1) {
        FDS_CK(c, launch_batch_jreduce(c->jnode, c->plan.nnode, c->gl0 / kLeafSites, c->slot / kLeafSites,
                                       c->n / kLeafSites, 2 * c->M, c->nb, nsteps,
                                       c->jsum + (size_t)step0 * c->nb * c->M * 2, c->stream, &c->ls));
        return 0;
    }
    const size_t row = (size_t)c->plan.nnode * c->M * 2;
    (void)row;
    FDS_CK(c, launch_tree_reduce_batched(c->jnode, c->plan.nnode, 2 * c->M, nsteps,
                                         c->jsum + (size_t)step0 * c->M * 2, c->jnode_scratch, c->stream, &c->ls));
    return 0;
}

// arguments of one single-column step P(from) -> P(to) with `valid` halo steps left; a batch is
// swept as one stacked array (the gaps hold the ghost cells)
static void primal_step_args(fds_ctx* c, StepArgs& a, int from, int to, long long valid, double s) {
    const double F = s + 8.0;
    a.X = c->P_all(from); a.Y = c->P_all(to); a.M = 1; a.n = c->nall;
    if (c->nb > 1) { a.lo = 0; a.hi = c->nall; a.Fpp = c->Fpp; a.slot = c->slot; }
    else { a.lo = -8 * (valid - 1); a.hi = c->n + 4 * (valid - 1); }
    a.F0 = F; a.Flast = F; a.dt = c->cfg.dt;
    a.exact = c->cfg.math == FDS_MATH_EXACT; a.jpart = nullptr; a.nleaf = c->nleaf;
}

// batch: (F0, Flast) of every problem for parameter s + ds[p] (and s + ds[p] + eps for the last column)
static int upload_forcing(fds_ctx* c, double s, double eps) {
    if (c->nb <= 1) return 0;
    for (int p = 0; p < c->nb; ++p) {
        const double sp = s + c->ds[p];
        c->Fhost[p] = make_double2(sp + 8.0, (sp + eps) + 8.0);
    }
    // pageable source: the copy is staged before the call returns, so Fhost may be rewritten afterwards
    FDS_CK(c, cudaMemcpyAsync(c->Fpp, c->Fhost.data(), (size_t)c->nb * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int fds_set_batch_params(fds_ctx* c, const double* ds) {
    for (int p = 0; p < c->nb; ++p) c->ds[p] = ds ? ds[p] : 0.0;
    return 0;
}

int fds_batch_count(const fds_ctx* c) { return c->nb; }

int fds_segment_sensitivities(fds_ctx* c, int64_t steps, double eps, double* J0_host, double* G_host) {
    FDS_ON_DEVICE(c);
    if (steps > c->jsum_cap || steps < 2) return fail(c, "fds_segment_sensitivities: needs 2 <= steps <= steps run");
    const size_t nj = (size_t)c->nb * steps * 2, ng = (size_t)c->nb * c->MO * 2;
    if (c->sens_cap < nj + ng) {
        cudaFree(c->sens);
        c->sens = nullptr;
        FDS_CK(c, cudaMalloc(&c->sens, (nj + ng) * sizeof(double)));
        c->sens_cap = nj + ng;
    }
    FDS_CK(c, launch_batch_sens(c->jsum, c->nb, c->M, steps, (double)c->cfg.n_global, eps, c->sens, c->sens + nj,
                                c->stream, &c->ls));
    FDS_CK(c, cudaMemcpyAsync(J0_host, c->sens, nj * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaMemcpyAsync(G_host, c->sens + nj, ng * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_primal_advance(fds_ctx* c, double s, int64_t step0, int64_t nsteps, int record) {
    FDS_ON_DEVICE(c);
    if (int r = sync_primal_from_ensemble(c)) return r;
    if (record) if (int r = ensure_jsum(c, step0 + nsteps)) return r;
    if (c->host_sys) return fail(c, "host-solver system: the solver runs on the host (fds_set_state / fds_set_ensemble)");
    if (c->toy) {
        FDS_CK(c, launch_toy_advance(c->cfg.system, c->cfg.math == FDS_MATH_EXACT, c->P(0), c->P(0), 1, nsteps, s, s,
                                     record ? c->jsum + (size_t)step0 * 2 : nullptr, c->stream, &c->ls));
        return 0;
    }
    if (c->nb > 1 && record) return fail(c, "fds_primal_advance: objective recording is not available for a batch");
    if (int r = upload_forcing(c, s, 0.0)) return r;
    const bool euler = c->cfg.system == FDS_SYS_L96_EULER;
    // primal marches of a single problem (run-up; the plugin called as u1, J = run(u0, s, steps)): up to
    // kMaxDilationOrder steps per pass in registers (the fused time-dilation march writing the state and, when J is
    // wanted, the leaf sums of every step); FDS_PRIMAL_FUSED=0 keeps the one-step kernel
    static const bool fused_ok = [] { const char* e = getenv("FDS_PRIMAL_FUSED"); return !(e && e[0] == '0'); }();
    for (long long j = 0; j < nsteps;) {
        if (c->pri_halo_valid < 1) return fail(c, "fds_primal_advance: primal halo exhausted; refresh it first");
        long long k = 1;
        if (fused_ok) {
            const long long v = c->pri_halo_valid;
            k = std::min<long long>(std::min<long long>(nsteps - j, v), kMaxDilationOrder);
            // recorded: the leaf sums of the k steps share the leaf buffer sized for one step of the M-column ensemble
            if (record) k = std::min<long long>(k, c->M);
            const long long nl = (c->n + 63) / 64;
            if (c->nb > 1)       // every stacked site, ghost cells included: what k passes of the one-step kernel leave valid
                FDS_CK(c, launch_primal_multi(c->P_all(0), c->P_all(1), 0, c->nall, (int)k, s + 8.0, c->cfg.dt,
                                              c->cfg.math == FDS_MATH_EXACT, euler, nullptr, 0, 0, c->Fpp, c->slot,
                                              c->stream, &c->ls));
            else
            FDS_CK(c, launch_primal_multi(c->P_all(0), c->P_all(1), -8 * (v - k), c->n + 4 * (v - k), (int)k, s + 8.0,
                                          c->cfg.dt, c->cfg.math == FDS_MATH_EXACT, euler, record ? c->jpart : nullptr,
                                          c->n, nl, nullptr, 1, c->stream, &c->ls));
            if (record)
                FDS_CK(c, launch_tree_reduce_batched(c->jpart, nl, 2, k, c->jsum + (size_t)(step0 + j) * 2, c->jscratch,
                                                     c->stream, &c->ls));
        } else {
            // the one-thread-per-element kernel (single column), one step
            StepArgs a;
            primal_step_args(c, a, 0, 1, c->pri_halo_valid, s);
            FDS_CK(c, launch_l96_simple(a, euler, c->stream, &c->ls));
            if (record) {
                FDS_CK(c, launch_stats_leaf(c->P(1), 1, c->n, c->jpart, c->nleaf, c->stream, &c->ls));
                FDS_CK(c, launch_tree_reduce(c->jpart, c->nleaf, 2, c->jsum + (size_t)(step0 + j) * 2, c->jscratch, c->stream, &c->ls));
            }
        }
        std::swap(c->Pbase[0], c->Pbase[1]);
        c->primal_synced = false;
        c->pri_halo_valid -= k;
        j += k;
    }
    return 0;
}

int fds_run(fds_ctx* c, double s, int64_t steps, double* Jsum_host) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world != 1 && !c->comm)
        return fail(c, "fds_run: world > 1, attach a communicator (fds_comm_init) or drive fds_primal_advance with halo exchanges");
    if (int r = sync_primal_from_ensemble(c)) return r;
    for (long long t = 0; t < steps;) {
        const long long hs = fds_halo_steps(c, steps - t);
        if (int r = fds_halo_wrap_primal(c, hs)) return r;
        if (int r = fds_primal_advance(c, s, t, hs, Jsum_host != nullptr)) return r;
        t += hs;
    }
    if (Jsum_host && steps > 0) if (int r = reduce_jsum_ranks(c, (size_t)steps * 2)) return r;
    if (Jsum_host && steps > 0)
        FDS_CK(c, cudaMemcpyAsync(Jsum_host, c->jsum, (size_t)steps * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int fds_segment_advance(fds_ctx* c, double s, double eps, int64_t step0, int64_t nsteps) {
    FDS_ON_DEVICE(c);
    if (c->host_sys) return fail(c, "host-solver system: advance the ensemble on the host (fds_get_ensemble / fds_set_ensemble)");
    if (int r = ensure_jsum(c, step0 + nsteps)) return r;
    if (c->toy) {
        // all nsteps in one launch, in place; column m+1 runs at parameter + epsilon (fds/segment.py:64)
        FDS_CK(c, launch_toy_advance(c->cfg.system, c->cfg.math == FDS_MATH_EXACT, c->Ecur(), c->Ecur(), c->M, nsteps,
                                     s, s + eps, c->jsum + (size_t)step0 * c->M * 2, c->stream, &c->ls));
        c->state_in_ensemble = true;
        c->primal_synced = false;
        c->tangents_implicit = true;
        c->eps_last = eps;
        return 0;
    }
    const double F0 = s + 8.0, Flast = (s + eps) + 8.0;   // fds/segment.py:64: parameter + epsilon
    if (nsteps > c->H) return fail(c, "fds_segment_advance: at most fds_halo_steps() steps per call");
    if (int r = upload_forcing(c, s, eps)) return r;
    static const bool emit_env = [] { const char* e = getenv("FDS_EMIT_PRIMAL"); return !(e && e[0] == '0'); }();
    const bool emit_ok = emit_env && c->stream_ok && c->cfg.system == FDS_SYS_L96_RK4;
    bool emitted = false;
    c->primal_synced = false;
    for (long long j = 0; j < nsteps;) {
        if (c->ens_halo_valid < 1) return fail(c, "fds_segment_advance: ensemble halo exhausted; refresh it first");
        // steps per pass: 4 (forward Euler: four chained pipes fit the register file), 2 (RK4), 1 (remainder)
        int k = 1;
        if (c->fuse2) {
            if (c->fuse4 && nsteps - j >= 4 && c->ens_halo_valid >= 4) k = 4;
            else if (nsteps - j >= 2 && c->ens_halo_valid >= 2) k = 2;
        }
        // the last pass of the call (RK4, two steps) also writes column 0 to the primal array: the next boundary
        // finds the state there instead of extracting it from the ensemble with a strided pass
        const bool emit = emit_ok && k == 2 && j + k >= nsteps;
        if (int r = one_step(c, c->E_all(c->cur), c->E_all(c->cur ^ 1), c->M, F0, Flast, c->ens_halo_valid, step0 + j,
                             j, k, emit ? c->P_all(0) : nullptr)) return r;
        emitted = emit;
        c->cur ^= 1;
        c->ens_halo_valid -= k;
        j += k;
    }
    if (int r = reduce_nodes(c, step0, nsteps)) return r;
    c->state_in_ensemble = true;
    c->tangents_implicit = true;
    c->primal_synced = emitted;
    if (emitted) c->pri_halo_valid = 0;
    c->eps_last = eps;
    return 0;
}

int fds_segment_jsums(fds_ctx* c, int64_t steps, double* out) {
    FDS_ON_DEVICE(c);
    if (steps > c->jsum_cap) return fail(c, "fds_segment_jsums: more steps than were run");
    FDS_CK(c, cudaMemcpyAsync(out, c->jsum, (size_t)steps * c->nb * c->M * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// the whole segment on the stream: halo refreshes (wrap / NCCL), the step kernels, the objective reduction
// (over the ranks too) and, if asked for, the copy of the sums to the host -- nothing is waited for
int fds_segment_enqueue(fds_ctx* c, double s, double eps, int64_t steps, double* Jsum_host) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world != 1 && !c->comm)
        return fail(c, "fds_segment: world > 1, attach a communicator (fds_comm_init) or drive fds_segment_advance with halo exchanges");
    for (long long t = 0; t < steps;) {
        const long long hs = fds_halo_steps(c, steps - t);
        if (int r = fds_halo_wrap_ensemble(c, hs)) return r;
        if (int r = fds_segment_advance(c, s, eps, t, hs)) return r;
        t += hs;
    }
    const size_t len = (size_t)steps * c->nb * c->M * 2;
    if (int r = reduce_jsum_ranks(c, len)) return r;
    if (Jsum_host && steps > 0)
        FDS_CK(c, cudaMemcpyAsync(Jsum_host, c->jsum, len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

int fds_segment(fds_ctx* c, double s, double eps, int64_t steps, double* Jsum_host) {
    FDS_ON_DEVICE(c);
    if (int r = fds_segment_enqueue(c, s, eps, steps, Jsum_host)) return r;
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------- boundary --------------------------------- //
int fds_boundary_begin(fds_ctx* c, int from_ensemble) {
    FDS_ON_DEVICE(c);
    if (from_ensemble) {
        c->state_in_ensemble = true;   // column 0 of the ensemble is the state of record
        if (int r = sync_primal_from_ensemble(c)) return r;
    } else if (c->state_in_ensemble) {
        if (int r = sync_primal_from_ensemble(c)) return r;
    }
    return 0;
}

int fds_boundary_dxdt(fds_ctx* c, double s, int from_ensemble, double eps) {
    FDS_ON_DEVICE(c);
    if (c->dil_mode == FDS_DILATION_GIVEN) {
        // d was put there by fds_set_dxdt (run_ddt callable of the reference, fds/timedilation.py:54-61)
    } else if (c->dil_mode == FDS_DILATION_ANALYTIC) {
        if (c->pri_halo_valid < 1) return fail(c, "fds_boundary_dxdt: primal halo must be valid");
        if (int r = upload_forcing(c, s, 0.0)) return r;
        FDS_CK(c, launch_l96_ddt(c->P_all(0), c->d_alloc, c->nall, s + 8.0, c->cfg.dt, c->cfg.math == FDS_MATH_EXACT,
                                 c->cfg.system == FDS_SYS_L96_EULER, c->nb > 1 ? c->Fpp : nullptr, c->slot,
                                 c->stream, &c->ls));
        c->pri_halo_valid = 0;
    } else if (c->time_dilation && c->host_sys) {
        // u1, u2, u3 were put into P(1..3) by fds_set_primal_history
        FDS_CK(c, launch_dxdt(c->P(0), c->P(1), c->P(2), c->P(3), c->c4, c->d, c->n, c->stream, &c->ls));
    } else if (c->time_dilation && c->toy) {
        for (int j = 1; j <= 3; ++j)
            FDS_CK(c, launch_toy_advance(c->cfg.system, c->cfg.math == FDS_MATH_EXACT, c->P(j - 1), c->P(j), 1, 1, s, s,
                                         nullptr, c->stream, &c->ls));
        FDS_CK(c, launch_dxdt(c->P(0), c->P(1), c->P(2), c->P(3), c->c4, c->d, c->n, c->stream, &c->ls));
    } else if (c->time_dilation) {
        if (c->pri_halo_valid < c->stencil.order)
            return fail(c, "fds_boundary_dxdt: primal halo must be valid for " + std::to_string(c->stencil.order) + " steps");
        if (int r = upload_forcing(c, s, 0.0)) return r;
        const bool split = getenv("FDS_DILATION_SPLIT") != nullptr;            // cross-check path: one kernel per step
        if (split && c->stencil.order == 3) {
            long long valid = 3;
            for (int j = 1; j <= 3; ++j) {
                StepArgs a;
                primal_step_args(c, a, j - 1, j, valid, s);
                FDS_CK(c, launch_l96_simple(a, c->cfg.system == FDS_SYS_L96_EULER, c->stream, &c->ls));
                --valid;
            }
            FDS_CK(c, launch_dxdt(c->P_all(0), c->P_all(1), c->P_all(2), c->P_all(3), c->c4, c->d_alloc, c->nall, c->stream, &c->ls));
        } else {
            // u_1 .. u_order and d = sum c_k u_k in one pass, nothing but d written
            FDS_CK(c, launch_dilation_fused(c->P_all(0), c->d_alloc, c->nall, c->stencil, s + 8.0, c->cfg.dt,
                                            c->cfg.math == FDS_MATH_EXACT, c->cfg.system == FDS_SYS_L96_EULER,
                                            c->nb > 1 ? c->Fpp : nullptr, c->slot, c->stream, &c->ls));
        }
        c->pri_halo_valid = 0;
    } else {
        FDS_CK(c, cudaMemsetAsync(c->d_alloc, 0, (size_t)c->nall * sizeof(double), c->stream));
    }
    const int src = from_ensemble ? SRC_ENSEMBLE : SRC_TANGENT;
    if (!from_ensemble) if (int r = materialize_tangents(c)) return r;
    FDS_CK(c, do_gram(c, src, from_ensemble ? 1.0 / eps : 0.0));
    return 0;
}

int fds_boundary_factor(fds_ctx* c, int do_qr, int from_ensemble, double eps) {
    FDS_ON_DEVICE(c);
    c->qr_single = false;
    if (do_qr && (c->fast_boundary || c->mma_boundary) && c->qr_mode == 0) {
        // one Cholesky pass loses orthogonality like cond^2 u: take it when that is below 2e-13
        FDS_CK(c, launch_small_single(c->small, c->m, c->time_dilation ? 1 : 0, c->stream, &c->ls, c->batch()));
        const std::string keep = c->err;
        const int r = check_small_status(c, "CholeskyQR");
        if (r == 0 && c->cond_last * c->cond_last * 2.3e-16 < kSinglePassTol) {
            c->qr_single = true;
            c->passes_last = 1;
            c->pend_src = from_ensemble ? SRC_ENSEMBLE : SRC_TANGENT;
            c->pend_inv_eps = from_ensemble ? 1.0 / eps : 0.0;
            return 0;
        }
        if (r == 1) return r;             // CUDA error
        c->err = keep;                    // ill conditioned or rank deficient: CholeskyQR2 decides
    }
    FDS_CK(c, launch_small1(c->small, c->m, c->time_dilation ? 1 : 0, do_qr, c->stream, &c->ls, c->batch()));
    const int src = from_ensemble ? SRC_ENSEMBLE : SRC_TANGENT;
    bool shifted = false;
    if (do_qr || c->fast_boundary) {
        const std::string keep = c->err;
        const int r = check_small_status(c, "CholeskyQR pass 1");
        // (the bound is garbage-in-garbage-out for a numerically singular Gram: anything that is not a number below
        // the limit counts, and so does a failed pivot)
        const bool too_ill = do_qr && r != 1 && (r == 2 || !(c->cond_last < kShiftedQrCond));
        if (too_ill && (c->cfg.world == 1 || c->comm)) {
            // The Gram of the tangent block is numerically singular (cond_2 of the block beyond ~1e8: a subspace that
            // reaches far into the stable directions, e.g. m = N with a wide Lyapunov spectrum -- the reference's
            // Householder QR, pascal_lite/operators/linalg.py:21, does not care).  Shifted CholeskyQR3: a first
            // factor of G + sigma I, sigma = 11 (m n + m (m + 1)) u trace(G) >= the bound of Fukaya et al. (2020),
            // always exists and leaves a block with cond_2 <~ 1e7, on which CholeskyQR2 (middle pass + pass 2) is
            // accurate; R = (L1s L2 L3)^T.
            c->err = keep;
            const double nn = (double)c->cfg.n_global, mm = (double)c->m;
            const double shift_rel = 11.0 * (mm * nn + mm * (mm + 1.0)) * 1.1102230246251565e-16;
            FDS_CK(c, launch_small1(c->small, c->m, c->time_dilation ? 1 : 0, 1, c->stream, &c->ls, c->batch(), shift_rel));
            if (int r2 = check_small_status(c, "shifted CholeskyQR3 pass 1")) return r2;
            FDS_CK(c, do_apply(c, src, from_ensemble ? 1.0 / eps : 0.0, c->T, nullptr, 0.0));
            FDS_CK(c, do_gram(c, SRC_TANGENT, 0.0));
            if (int r2 = allreduce_gram(c, c->small.gram)) return r2;
            FDS_CK(c, launch_small_mid(c->small, c->m, c->stream, &c->ls, c->batch()));
            if (int r2 = check_small_status(c, "shifted CholeskyQR3 pass 2")) return r2;
            FDS_CK(c, do_apply(c, SRC_TANGENT, 0.0, c->T, nullptr, 0.0));
            shifted = true;
        } else if (r != 0) {
            return r;
        }
    }
    if (!shifted) FDS_CK(c, do_apply(c, src, from_ensemble ? 1.0 / eps : 0.0, c->T, nullptr, 0.0));
    c->tangents_implicit = false;
    c->shifted_last = shifted;
    if (do_qr) FDS_CK(c, do_gram(c, SRC_TANGENT, 0.0));
    return 0;
}

int fds_boundary_finish(fds_ctx* c, int do_qr, double eps, int build) {
    FDS_ON_DEVICE(c);
    const bool ens = (build & FDS_BUILD_ENSEMBLE) != 0;
    // explicit tangents are written unless the caller only wants the next ensemble (hot loop)
    const bool keepT = !ens || (build & FDS_BUILD_KEEP_TANGENTS) != 0 || !(c->fast_boundary || c->mma_boundary);
    c->tangents_implicit = false;
    if (do_qr && c->qr_single) {
        c->qr_single = false;
        FDS_CK(c, do_apply(c, c->pend_src, c->pend_inv_eps, keepT ? c->T : nullptr, ens ? c->Ecur() : nullptr, eps));
        if (!keepT) c->tangents_implicit = true;    // the new ensemble is the record of the tangents
    } else if (do_qr) {
        c->passes_last = c->shifted_last ? 3 : 2;
        FDS_CK(c, launch_small2(c->small, c->m, c->stream, &c->ls, c->batch()));
        if (int r = check_small_status(c, "CholeskyQR pass 2")) return r;
        FDS_CK(c, do_apply(c, SRC_TANGENT, 0.0, keepT ? c->T : nullptr, ens ? c->Ecur() : nullptr, eps));
        if (!keepT) c->tangents_implicit = true;    // T holds the pass-1 intermediate; the ensemble is the record
    } else if (ens) {
        FDS_CK(c, launch_make_ensemble(c->P_all(0), c->T_alloc, c->M, eps, c->E_all(c->cur), c->nall, c->stream, &c->ls));
    }
    if (ens) { c->ens_halo_valid = 0; c->eps_last = eps; }
    c->state_in_ensemble = false;
    return 0;
}

int fds_boundary_results(fds_ctx* c, double* R, double* b, double* G_dil, double* g_dil) {
    FDS_ON_DEVICE(c);
    // batch: one row per problem -- R[nb][m][m], b[nb][m], G_dil[nb][m], g_dil[nb]
    const size_t pitch = (size_t)c->small_stride * sizeof(double);
    auto fetch = [&](double* dst, const double* src, size_t count) {
        return cudaMemcpy2DAsync(dst, count * sizeof(double), src, pitch, count * sizeof(double), c->nb,
                                 cudaMemcpyDeviceToHost, c->stream);
    };
    if (R) FDS_CK(c, fetch(R, c->small.R, (size_t)c->m * c->m));
    if (b) FDS_CK(c, fetch(b, c->small.b, c->m));
    if (G_dil) FDS_CK(c, fetch(G_dil, c->small.gdil, c->m));
    if (g_dil) FDS_CK(c, fetch(g_dil, c->small.gdil + c->m, 1));
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- the device-gated boundary: no host round trip between its kernels ------------------------ //
// factor + finish of a boundary whose big passes are tensor-core kernels.  The adaptive CholeskyQR
// decision is taken by small_single_kernel on the device (two_pass flag next to the status word); the
// CholeskyQR2 kernels and both variants of the final apply are enqueued behind it, gated on that flag.
// With a communicator the second Gram is all-reduced unconditionally (9 KB; the ranks hold identical
// all-reduced first Grams, so they take the same branch).
static int gated_factor_finish(fds_ctx* c, int do_qr, int from_ensemble, double eps, int build) {
    const bool ens = (build & FDS_BUILD_ENSEMBLE) != 0;
    const bool keepT = !ens || (build & FDS_BUILD_KEEP_TANGENTS) != 0;
    const int src = from_ensemble ? SRC_ENSEMBLE : SRC_TANGENT;
    const double inv_eps = from_ensemble ? 1.0 / eps : 0.0;
    const int use_d = c->time_dilation ? 1 : 0;
    c->qr_single = false;
    if (!do_qr) {
        FDS_CK(c, launch_small1(c->small, c->m, use_d, 0, c->stream, &c->ls, c->batch()));
        FDS_CK(c, do_apply(c, src, inv_eps, c->T, nullptr, 0.0));
        c->tangents_implicit = false;
        if (ens) FDS_CK(c, launch_make_ensemble(c->P_all(0), c->T_alloc, c->M, eps, c->E_all(c->cur), c->nall, c->stream, &c->ls));
    } else {
        // every problem's own gate (the int next to its status word): small_single_kernel sets it to 0 or 1, so a
        // well-conditioned problem of a sweep skips the CholeskyQR2 kernels that an ill-conditioned one needs
        int* flag = c->small.status + 1;
        const long long fstride = 2 * c->small_stride;        // ints between two problems' status blocks
        FDS_CK(c, launch_small_single(c->small, c->m, use_d, c->stream, &c->ls, c->batch(), flag, c->qr_mode == 2, fstride));
        BatchDesc two = c->bd, one = c->bd;
        two.gate = flag; two.gate_want = 1; two.gate_stride = fstride;
        one.gate = flag; one.gate_want = 0; one.gate_stride = fstride;
        // CholeskyQR2 (gate = 1): pass 1 -> T, Gram of T -> gram2, pass 2
        FDS_CK(c, launch_small1(c->small, c->m, use_d, 1, c->stream, &c->ls, &two));
        FDS_CK(c, launch_apply_mma(src, c->Ecur(), c->T, c->d, c->small.A, c->m, inv_eps, c->n, c->T, nullptr, c->P(0),
                                   0.0, c->num_sms, c->stream, &c->ls, &two));
        FDS_CK(c, launch_gram_mma(SRC_TANGENT, c->Ecur(), c->T, c->d, c->m, 0.0, c->n, c->gram_part, c->max_part,
                                  c->gram2, c->num_sms, c->stream, &c->ls, &two));
        if (int r = allreduce_gram(c, c->gram2)) return r;
        FDS_CK(c, launch_small2(c->small, c->m, c->stream, &c->ls, &two, c->gram2));
        // the last pass: from the source with the single-pass factor (gate = 0), or from T with the second factor
        double* Tout = keepT ? c->T : nullptr;
        double* Eout = ens ? c->Ecur() : nullptr;
        FDS_CK(c, launch_apply_mma(src, c->Ecur(), c->T, c->d, c->small.A, c->m, inv_eps, c->n, Tout, Eout, c->P(0), eps,
                                   c->num_sms, c->stream, &c->ls, &one));
        FDS_CK(c, launch_apply_mma(SRC_TANGENT, c->Ecur(), c->T, c->d, c->small.A, c->m, 0.0, c->n, Tout, Eout, c->P(0), eps,
                                   c->num_sms, c->stream, &c->ls, &two));
        c->tangents_implicit = !keepT;            // without T the new ensemble is the record of the tangents
    }
    if (ens) { c->ens_halo_valid = 0; c->eps_last = eps; }
    c->state_in_ensemble = false;
    return 0;
}

int64_t fds_boundary_out_doubles(const fds_ctx* c) { return (int64_t)c->m * c->m + 2 * c->m + 3; }

// layout of one problem's block: R[m*m], b[m], G_dil[m], g_dil, {int32 status, int32 two_pass}, cond
static int decode_boundary_block(fds_ctx* c, const double* blk, int problem) {
    const size_t o = (size_t)c->m * c->m + 2 * c->m + 1;
    int st[2];
    std::memcpy(st, blk + o, sizeof(st));
    if (st[0] != 0) {
        const char* what = st[0] >= kSmallPass2 ? "CholeskyQR pass 2" : (st[0] >= kSmallPass1 ? "CholeskyQR pass 1" : "CholeskyQR");
        std::string msg = std::string(what) + ": Cholesky pivot <= 0 at column " + std::to_string(st[0] % 1000);
        if (c->nb > 1) msg += " of problem " + std::to_string(problem);
        return fail(c, msg + " (tangent block numerically rank deficient)");
    }
    return 0;
}

int fds_boundary_check(fds_ctx* c, const double* out, int do_qr) {
    if (!do_qr) return 0;
    const size_t len = (size_t)fds_boundary_out_doubles(c);
    double cond = 0.0;
    int two = 0;
    for (int p = 0; p < c->nb; ++p) {
        if (int r = decode_boundary_block(c, out + p * len, p)) return r;
        cond = std::max(cond, out[p * len + len - 1]);
        int st[2];
        std::memcpy(st, out + p * len + len - 2, sizeof(st));
        two |= st[1];
    }
    c->cond_last = cond;
    if (c->gated) c->passes_last = two ? 2 : 1;
    return 0;
}

// One whole boundary enqueued on the stream (world > 1: with the context's communicator).  `out` receives
// fds_boundary_out_doubles() doubles per problem once the stream has reached this point (page-locked memory
// keeps the copy asynchronous); check it with fds_boundary_check() after fds_sync().  Contexts whose boundary
// cannot run gated (m other than 16 / 32, small ODE systems) do the same work with their host round trips.
int fds_boundary_enqueue(fds_ctx* c, double s, double eps, int from_ensemble, int do_qr, int build, int64_t halo_steps,
                         double* out) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world != 1 && !c->comm)
        return fail(c, "fds_boundary: world > 1, attach a communicator (fds_comm_init) or drive the phases with reductions in between");
    if (int r = fds_boundary_begin(c, from_ensemble)) return r;
    if (c->time_dilation) if (int r = fds_halo_wrap_primal(c, c->stencil.order)) return r;
    if (int r = fds_boundary_dxdt(c, s, from_ensemble, eps)) return r;
    if (int r = allreduce_gram(c, c->small.gram)) return r;
    const size_t len = (size_t)fds_boundary_out_doubles(c);
    if (c->gated) {
        if (int r = gated_factor_finish(c, do_qr, from_ensemble, eps, build)) return r;
    } else {
        if (int r = fds_boundary_factor(c, do_qr, from_ensemble, eps)) return r;
        if (do_qr && !c->qr_single) if (int r = allreduce_gram(c, c->small.gram)) return r;
        if (int r = fds_boundary_finish(c, do_qr, eps, build)) return r;
    }
    if (build && halo_steps > 0) if (int r = fds_halo_wrap_ensemble(c, halo_steps)) return r;
    if (out != nullptr)
        FDS_CK(c, cudaMemcpy2DAsync(out, len * sizeof(double), c->small.R, (size_t)c->small_stride * sizeof(double),
                                    len * sizeof(double), c->nb, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

int fds_boundary(fds_ctx* c, double s, double eps, int from_ensemble, int do_qr, int build, int64_t halo_steps,
                 double* R, double* b, double* G_dil, double* g_dil) {
    FDS_ON_DEVICE(c);
    if (int r = fds_boundary_enqueue(c, s, eps, from_ensemble, do_qr, build, halo_steps, c->out_host)) return r;
    FDS_CK(c, cudaStreamSynchronize(c->stream));
    if (int r = fds_boundary_check(c, c->out_host, do_qr)) return r;
    const size_t len = (size_t)fds_boundary_out_doubles(c), mm = (size_t)c->m * c->m, m = (size_t)c->m;
    for (int p = 0; p < c->nb; ++p) {
        const double* blk = c->out_host + p * len;
        if (do_qr && R) std::memcpy(R + p * mm, blk, mm * sizeof(double));
        if (do_qr && b) std::memcpy(b + p * m, blk + mm, m * sizeof(double));
        if (G_dil) std::memcpy(G_dil + p * m, blk + mm + m, m * sizeof(double));
        if (g_dil) g_dil[p] = blk[mm + 2 * m];
    }
    return 0;
}

int fds_run_segment_host(fds_ctx* c, double* u, double* V, double* v, double s, double eps, int64_t steps,
                         double* Jsum_host) {
    FDS_ON_DEVICE(c);
    if (c->cfg.world != 1) return fail(c, "fds_run_segment_host: world == 1 only");
    if (int r = fds_set_state(c, u)) return r;
    if (int r = fds_set_tangents(c, V, v)) return r;
    FDS_CK(c, launch_make_ensemble(c->P_all(0), c->T_alloc, c->M, eps, c->E_all(c->cur), c->nall, c->stream, &c->ls));
    c->eps_last = eps;
    if (int r = fds_segment(c, s, eps, steps, Jsum_host)) return r;
    if (int r = fds_get_state(c, u)) return r;
    return fds_get_tangents(c, V, v);
}

// ---------------------------------- LSS solve -------------------------------- //
int fds_lss_solve(int device, const double* R, const double* b, int64_t K, int m, double* alpha) {
    return fds_lss_solve_batch(device, R, b, K, m, 1, alpha);
}

// Device buffers of the context-free LSS solve, kept between calls (one set per GPU, grown on demand): with
// verbose=True the reference solves after EVERY segment (fds/fds.py:161), which used to cost five cudaMalloc and
// four blocking copies per call.  One lock: concurrent solves on the same process are serialised.
namespace {
struct LssScratch {
    double* buf = nullptr;      // [R | b | alpha | work]
    int* status = nullptr;
    size_t cap = 0, cap_st = 0;
    cudaStream_t stream = nullptr;
};
std::mutex g_lss_mutex;
LssScratch g_lss[64];
}  // namespace

int fds_lss_solve_batch(int device, const double* R, const double* b, int64_t K, int m, int64_t nbatch, double* alpha) {
    if (K < 1 || m < 1 || nbatch < 1 || lss_smem_bytes(m) > 220 * 1024) { g_create_err = "fds_lss_solve: bad K, m or batch"; return 2; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e == cudaSuccess && (device < 0 || device >= ndev || device >= 64)) e = cudaErrorInvalidDevice;
    if (e != cudaSuccess) { g_create_err = std::string("fds_lss_solve: ") + cudaGetErrorString(e) + " -- no CPU fallback"; return 3; }
    DeviceGuard guard__(device);
    std::lock_guard<std::mutex> lock(g_lss_mutex);
    LssScratch& sc = g_lss[device];
    const size_t mm = (size_t)m * m, nbz = (size_t)nbatch;
    const size_t nR = nbz * K * mm, nb_ = nbz * K * m, nw = nbz * ((size_t)2 * K * mm + (size_t)2 * K * m);
    const size_t need = nR + 2 * nb_ + nw;
#define CKL(call)                                                                          \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) { g_create_err = std::string(#call) + ": " + cudaGetErrorString(e__); return 1; } \
    } while (0)
    if (!sc.stream) CKL(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
    if (sc.cap < need) {
        cudaFree(sc.buf); sc.buf = nullptr; sc.cap = 0;
        CKL(cudaMalloc(&sc.buf, need * sizeof(double)));
        sc.cap = need;
    }
    if (sc.cap_st < nbz) {
        cudaFree(sc.status); sc.status = nullptr; sc.cap_st = 0;
        CKL(cudaMalloc(&sc.status, nbz * sizeof(int)));
        sc.cap_st = nbz;
    }
    double *dR = sc.buf, *db = dR + nR, *da = db + nb_, *dw = da + nb_;
    std::vector<int> sts(nbz, 0);
    CKL(cudaMemcpyAsync(dR, R, nR * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
    CKL(cudaMemcpyAsync(db, b, nb_ * sizeof(double), cudaMemcpyHostToDevice, sc.stream));
    CKL(launch_lss_solve(dR, db, K, m, da, dw, sc.status, sc.stream, nullptr, (int)nbatch));
    CKL(cudaMemcpyAsync(alpha, da, nb_ * sizeof(double), cudaMemcpyDeviceToHost, sc.stream));
    CKL(cudaMemcpyAsync(sts.data(), sc.status, nbz * sizeof(int), cudaMemcpyDeviceToHost, sc.stream));
    CKL(cudaStreamSynchronize(sc.stream));
#undef CKL
    for (size_t p = 0; p < nbz; ++p)
        if (sts[p] != 0) {
            g_create_err = "fds_lss_solve: Schur complement not positive definite at row " + std::to_string(sts[p]) +
                           " (problem " + std::to_string(p) + ")";
            return 4;
        }
    return 0;
}

}  // extern "C"
