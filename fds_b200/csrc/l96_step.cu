// Lorenz-96 ensemble time-step kernels for sm_100a.
//
// K2 of SURVEY.md 2.3: replaces the m+2 pooled solver runs of one segment
// (reference fds/segment.py:56-64; solver body tests/solvers/mock_fun3d/fun3d:48-57).
//
// l96_rk4_stream_kernel  -- the hot kernel.  Site-major ensemble [site][M] in HBM.
//   Each CTA owns R strips of consecutive sites; thread (r, k) marches along strip r
//   for ensemble column k with the whole 4-stage RK4 software-pipelined in registers
//   (l96_march.cuh), so a site is read once and written once per step (16 B per
//   trajectory-site-step) and no intermediate stage ever touches shared or global
//   memory.  Input arrives through a ring of shared-memory stages filled by 1-D TMA
//   bulk copies (cp.async.bulk + mbarrier complete_tx) issued by a producer warp;
//   every strip is a contiguous chunk of the site-major array, so each copy is one
//   contiguous 8*M*8-byte transfer.  Output goes straight from registers to HBM,
//   coalesced across the columns of a site.  Objective sums (sum x, sum x^2 per
//   column) are accumulated on the fly in a specified binary tree.
#include <algorithm>
#include <cstdlib>
#include "kernels.h"
#include "l96_march.cuh"
#include "l96_march2.cuh"
#include "l96_marchk.cuh"

namespace fds {

// ------------------------------ streaming kernel ----------------------------- //
struct StreamParams {
    const double* X;
    double* Y;
    double* jpart;
    long long nnode;
    long long a0, ls, nstrips;
    long long n;
    int nblk, node_shift;
    int M, R, run_stride, nstages, nconsumer;  // nconsumer = R*M
    int pack;                                  // 1: (strip, column) pairs are packed over ALL consumer lanes of the grid: a strip may
                                               // straddle two CTAs, which both stage its input; R = stage slots per CTA
    int dedicated;                             // 1: the last warp only produces; 0: warp 0 produces in between its blocks
    double F0, Flast;
    double* jpart2;                            // two steps per pass: node partials of the second step
    long long jstride;                         // K steps per pass: step j writes its node partials at jpart + j * jstride
    int nfuse;                                 // time steps per pass (1, 2 or 4)
    int fuse2;                                 // 1: two steps per pass (l96_march2.cuh)
    const double2* Fpp;                        // per-problem (F0, Flast) of a batch, or nullptr
    double* P0out;                             // EMIT kernels: column 0 of the output goes here as well (site 0)
    unsigned long long* cycles;                // profiling: sum of the CTAs' clock64() spans, number of CTAs (or nullptr)
    int spp;                                   // strips per problem (batch)
    int prefetch;                              // shared-producer mode: stages in flight ahead of the consumers (<= nstages - 2)
    int chunk;                                 // sites per strip and stage: 8 (one block), or 16 (two / four steps per pass: a
                                               // stage holds TWO blocks, so the ring is synchronised once per pair of blocks)
    Rk4Coef c;
};

constexpr int kMaxStages = 8;
constexpr uint32_t kPairStride = 256 * 8;   // bytes between two entries of a thread's pair-sum table (one row of 256 threads)
constexpr int kStreamHdrBytes = 256;      // full[8], empty[8], sink, padding: the per-thread tables start here

// The TMA producer role.  Dedicated mode: one extra warp does nothing else.  Shared mode (all warps
// compute): the consumer warps take turns -- at the end of consumer block c, warp (c mod nwarps)
// fetches block c + nstages - 2 into the stage that block c - 2 occupied (every warp has let go of it
// unless it lags two blocks behind), so no warp is systematically slower than the others.
// Lane r issues the bulk copy of strip r: one contiguous 8 * M * 8-byte chunk per strip and block.
struct StreamProducer {
    const StreamParams* P;   // kernel parameters (constant bank)
    const double* src;       // this lane's strip, first input site of block 0 (lane < R)
    double* dst0;            // this lane's slot in stage 0
    uint64_t* full;          // empty = full + kMaxStages
    int lane, nit, rcta;     // rcta: strips this CTA stages
    int it, st;              // next block to fetch and its stage
    uint32_t par;

    FDS_D void init(const StreamParams* P_, double* sdata, uint64_t* full_, int lane_, int first_off, int nit_) {
        P = P_; full = full_; lane = lane_; nit = nit_;
        const int R = P->R;
        rcta = R;
        // strips past the end of the plan (last CTA only) mirror the last strip
        long long strip = (long long)blockIdx.x * R + (lane < R ? lane : 0);
        if (P->pack) {
            const long long g0 = (long long)blockIdx.x * P->nconsumer;           // first (strip, column) pair of this CTA
            const long long sf = g0 / P->M;
            long long sl = (g0 + P->nconsumer - 1) / P->M;
            sl = sl < P->nstrips - 1 ? sl : P->nstrips - 1;
            rcta = sl >= sf ? (int)(sl - sf + 1) : 1;
            strip = sf + (lane < rcta ? lane : 0);
        }
        strip = strip < P->nstrips - 1 ? strip : P->nstrips - 1;
        src = P->X + (P->a0 + strip * P->ls + first_off) * P->M;
        dst0 = sdata + (lane < R ? lane : 0) * P->run_stride;
        it = 0; st = 0; par = 0;
    }
    FDS_D bool done() const { return it >= nit; }
    // fetch stage `it` (P->chunk input sites of every strip) into ring slot `st`
    FDS_D void issue() { issue_at(it, st, par); }
    FDS_D void issue_at(int it, int st, uint32_t par) {
        uint64_t* empty = full + kMaxStages;
        const int R = P->R, M = P->M, stage_stride = R * P->run_stride;
        const int CH = P->chunk;
        const uint32_t chunk_bytes = (uint32_t)(CH * M * sizeof(double));
        if (it >= P->nstages) mbar_wait(&empty[st], par ^ 1u);
        if (lane == 0) mbar_expect_tx(&full[st], chunk_bytes * (uint32_t)rcta);
        __syncwarp();
        if (lane < rcta) bulk_g2s(dst0 + (size_t)st * stage_stride, src + (long long)it * (CH * M), chunk_bytes, &full[st]);
        for (int r = lane + 32; r < rcta; r += 32) {         // more than 32 strips per CTA (tiny M; never packed)
            long long strip = (long long)blockIdx.x * R + r;
            strip = strip < P->nstrips - 1 ? strip : P->nstrips - 1;
            const long long off = (strip - ((long long)blockIdx.x * R + lane < P->nstrips - 1
                                                ? (long long)blockIdx.x * R + lane : P->nstrips - 1)) * P->ls;
            bulk_g2s(dst0 + (size_t)st * stage_stride + (r - lane) * P->run_stride,
                     src + (off + (long long)CH * it) * M, chunk_bytes, &full[st]);
        }
    }
    FDS_D void advance() {
        ++it;
        if (++st == P->nstages) { st = 0; par ^= 1u; }
    }
};

// MC = number of columns when known at compile time (every p*M offset becomes an immediate),
// 0 = use the run-time value
// MID = the march calls mid_block() once in EVERY block (the kernels that advance two or four steps per pass): the
// ring bookkeeping then happens there, inside straight-line code, and the block boundary carries only the barrier
// fallback, the stage release and the (rare) copy issue.
template <int MC, bool MID = false, bool EMIT = false>
struct StreamIO {
    const double* sdata;     // this thread's (run, column) slot in stage 0
    int stage_stride;        // doubles per stage
    int M;
    FDS_D int cols() const { return MC > 0 ? MC : M; }
    int nstages;
    uint32_t full_s;         // shared-window address of full[0]; empty[s] = full[kMaxStages + s]
    double* yrow;            // output cursor: site a + 8b - 3 of this thread's column
    double* prow;            // EMIT: the same site in the single-column primal array (column-0 lanes store there too)
    bool col0;
    double* jp;              // node 0, already offset by column (or nullptr)
    long long jstride;       // K steps per pass: step j's partials at jp + j * jstride
    int levels;              // node levels above the leaf (stride of the second step's counter storage)
    long long nnode;
    double* nsm;             // node-counter storage of this thread: nsm[(2*level + which) * nthr]
    uint32_t pls_s, plc_s;   // pair sums of the open leaves of this thread (l96_march2.cuh), shared-window address of
                             // entry (sum, slot) = pls_s + (4 * sum + slot) * kPairStride; plc_s: cursor on a slot
    int nthr;
    int lane;
    StreamProducer* prod;    // shared-producer mode: every consumer warp carries the producer state
    int warp, nwarps, turn;  // whose turn it is to fetch
    // ring position
    int st;
    uint32_t par;
    const double* cur;
    // MID only: a ring stage holds TWO 8-site blocks of every strip, so the ring is synchronised (barrier wait, stage
    // release, probe, copy duty) once per PAIR of blocks; the position is carried as shared-window addresses (no index
    // arithmetic, no generic->shared conversion in the loop).  H = which half of its stage a block is (0 / 1) when
    // known at compile time (the unchecked body: two blocks per iteration), -1: tracked in `half`.
    uint32_t cur_s, nxt_s;   // this thread's slot for the current block / for the next one (set by mid_block)
    uint32_t base_s;         // this thread's slot in the current stage (first half)
    uint32_t bar_s, bar_end; // `full` barrier of the current stage; one past the last stage's
    uint32_t stage_bytes, ring_bytes, half_bytes;
    uint32_t prev_empty;     // `empty` barrier of the previous stage, not yet released (first stage: a sink)
    int half;                // which half of its stage the next block to begin is (dynamic tracking)
    bool need_wait;          // the next block to begin opens a new stage
    int stg, ahead;          // stages begun so far; stages the copies run ahead of the consumers
    int nit;                 // stages per strip
    int duty;                // the next stage at whose end this warp issues copies (stage duty + ahead), INT_MAX: none left
    bool mine;               // this warp issues the copies of stage stg - 1 + ahead at the end of the block
    bool ready;              // the next stage's `full` phase was already seen complete by mid_block()

    template <int H = -1>
    FDS_D void begin_block(int) {
        if (MID) {
            if (H >= 0 ? H == 0 : need_wait) {
                if (!ready) mbar_wait_s(bar_s, par);
                // release the previous stage: EVERY lane arrives for itself (the barrier counts lanes), so no
                // warp-wide sync is needed -- fewer instructions between the FP64 streams of two blocks
                mbar_arrive_s(prev_empty);
                need_wait = false;
            }
            cur_s = nxt_s;
        } else {
            mbar_wait_s(full_s + 8u * (uint32_t)st, par);
            cur = sdata + (size_t)st * stage_stride;
        }
    }
    // the ring bookkeeping runs in the MIDDLE of a block, inside straight-line code where its latencies hide behind
    // the FP64 stream: in the second block of a stage the probe of the NEXT stage's `full` barrier, whose turn it
    // is to fetch, and the position of the next stage.  The producer position is a function of the consumer's
    // (stage c + `ahead` goes into slot st + ahead), so no separate producer state is carried.
    template <int H = -1>
    FDS_D void mid_block() {
        if (!MID) return;
        if (H >= 0 ? H == 1 : half == 1) {
            ++stg;
            prev_empty = bar_s + 8u * kMaxStages;
            bar_s += 8u;
            base_s += stage_bytes;
            if (bar_s == bar_end) {      // wrap: step back by the length of the ring
                bar_s -= 8u * (uint32_t)nstages;
                base_s -= ring_bytes;
                par ^= 1u;
            }
            nxt_s = base_s;
            // all lanes see the same phase: the vote makes the branch on it warp-uniform for the compiler as well
            ready = __all_sync(0xffffffffu, mbar_test_s(bar_s, par));
            need_wait = true;
        } else {
            // copy duty at the end of the FIRST block of this stage (every nwarps-th stage, while any is left): stage
            // stg + ahead is then three blocks away, and the slot it reuses was left by every warp a block ago
            mine = stg == duty;
            nxt_s = base_s + half_bytes;      // (the rest of THIS block still loads through cur_s)
        }
    }
    FDS_D int warp_max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    FDS_D int warp_min(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    FDS_D double load(int p) const {
        if (MID) return lds_f64(cur_s + (uint32_t)(p * cols()) * 8u);
        return cur[p * cols()];
    }
    FDS_D void store(int p, double v) {
        yrow[p * cols()] = v;
        if (EMIT && col0) prow[p] = v;
    }
    template <int H = -1>
    FDS_D void end_block(int) {
        if (MID) {
            if ((H >= 0 ? H == 0 : half == 0) && mine) {
                // this stage (stg) sits in the slot of bar_s with parity par; stage stg + ahead goes `ahead` slots on
                const int pst = (int)((bar_s - full_s) >> 3);
                const int q = pst + ahead;
                const bool w = q >= nstages;
                prod->issue_at(stg + ahead, w ? q - nstages : q, w ? par ^ 1u : par);
                duty += nwarps;
                if (duty + ahead >= nit) duty = 0x7fffffff;
            }
            if (H < 0) half ^= 1;
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive_s(full_s + 8u * (uint32_t)(kMaxStages + st));
            if (prod != nullptr) {
                if (!prod->done()) {
                    if (turn == warp) prod->issue();
                    prod->advance();
                }
                if (++turn == nwarps) turn = 0;
            }
            if (++st == nstages) { st = 0; par ^= 1u; }
        }
        yrow += 8 * cols();
        if (EMIT) prow += 8;
    }
    FDS_D void emit_node(long long node, double s1, double s2) {
        if (jp != nullptr && node >= 0 && node < nnode)
            *reinterpret_cast<double2*>(jp + node * (2LL * cols())) = make_double2(s1, s2);
    }
    FDS_D double ns_get(int l, int q) const { return nsm[(2 * l + q) * nthr]; }
    FDS_D void ns_put(int l, int q, double v) { nsm[(2 * l + q) * nthr] = v; }
    // two steps per pass: step 1 uses the first `levels` slots, step 2 the next
    FDS_D void emit_node2(int step, long long node, double s1, double s2) {
        double* j = jp != nullptr ? jp + step * jstride : nullptr;
        if (j != nullptr && node >= 0 && node < nnode)
            *reinterpret_cast<double2*>(j + node * (2LL * cols())) = make_double2(s1, s2);
    }
    FDS_D void pl_put(int sum, int slot, double v) { sts_f64(pls_s + (uint32_t)(4 * sum + slot) * kPairStride, v); }
    FDS_D double pl_get(int sum, int slot) const { return lds_f64(pls_s + (uint32_t)(4 * sum + slot) * kPairStride); }
    FDS_D void pl_cur_set(int slot) { plc_s = pls_s + (uint32_t)slot * kPairStride; }
    FDS_D void pl_cur_next() { plc_s += kPairStride; }
    template <int SUM>
    FDS_D void pl_cur_put(double v) { sts_f64_off<SUM * 4 * (int)kPairStride>(plc_s, v); }
    FDS_D double ns_get2(int step, int l, int q) const { return nsm[(2 * (step * levels + l) + q) * nthr]; }
    FDS_D void ns_put2(int step, int l, int q, double v) { nsm[(2 * (step * levels + l) + q) * nthr] = v; }
    FDS_D bool all(bool v) const { return __all_sync(0xffffffffu, v); }
};

// MODE 0: RK4, 1: forward Euler, 2: RK4, two steps per pass, 3: forward Euler, two steps per pass, 4: forward Euler, four
// steps per pass (l96_marchk.cuh) (<= 256 threads = 2 warps per scheduler: twice the pipeline state in up to 255 registers)
template <class MP, int MC, bool STATS, int MODE, bool EMIT = false>
__global__ void __launch_bounds__(MODE >= 2 ? 256 : 512, 1) l96_rk4_stream_kernel(const __grid_constant__ StreamParams p) {
    constexpr int KF = MODE == 4 ? 4 : (MODE >= 2 ? 2 : 1);     // time steps per pass
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    uint64_t* sink = full + 2 * kMaxStages;        // never waited on: takes the stage release of a strip's first block
    double* nsm = reinterpret_cast<double*>(smem_raw + kStreamHdrBytes);
    const int levels = p.node_shift - kLeafShift;
    double* pls = nsm + (size_t)(2 * KF) * levels * (blockDim.x - 32 * p.dedicated);
    // two / four steps per pass: 2 KF open leaves per thread, four pair sums each
    double* sdata = pls + (KF >= 2 ? (size_t)(8 * KF) * 256 : 0);

    const int tid = threadIdx.x;
    const int nwarps = (blockDim.x >> 5) - p.dedicated;   // consumer warps
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp-uniform by construction: say so
    const int stage_stride = p.R * p.run_stride;

    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MODE >= 2 ? nwarps * 32 : nwarps);      // two / four steps per pass: every lane arrives
        }
        mbar_init(sink, 0xfffff);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long clk0 = p.cycles != nullptr ? clock64() : 0;
    const long long last_strip = p.nstrips - 1;
    StreamProducer prod;
    // two / four steps per pass: the strip's first block is the SECOND half of stage 0 (the stage boundaries then fall
    // on even block indices, where the unchecked two-block body starts); the half before it is fetched, never read
    prod.init(&p, sdata, full, lane, FuseGeom<KF>::kInOff - (MODE >= 2 ? 8 : 0),
              MODE >= 2 ? (p.nblk + FuseGeom<KF>::kExtra + 1) / 2 : p.nblk + FuseGeom<KF>::kExtra);
    if (p.dedicated) {
        if (warp == nwarps) {
            while (!prod.done()) { prod.issue(); prod.advance(); }
            return;
        }
    } else {
        // prologue: warp 0 fills all but two stages; every warp tracks the producer position.  (Two-block stages:
        // stage 0 has no first block -- a strip starts in its second half -- so its copy duty is done here.)
        for (int j = 0; j < p.prefetch + (MODE >= 2 ? 1 : 0) && !prod.done(); ++j) {
            if (warp == 0) prod.issue();
            prod.advance();
        }
    }

    // -------------------- consumer threads: (run r, column k) ------------------ //
    // idle lanes of the last warp mirror the last thread, strips past the end of the plan mirror
    // the last strip: same loads, same arithmetic, same stores of the same values -- no
    // thread ever takes a path of its own
    const int t = tid < p.nconsumer ? tid : p.nconsumer - 1;
    int r = t / p.M, k = t - r * p.M;
    long long strip = (long long)blockIdx.x * p.R + r;
    if (p.pack) {
        // every lane of the grid carries its own (strip, column) pair; pairs past the end mirror the last one
        long long gidx = (long long)blockIdx.x * p.nconsumer + t;
        const long long total = p.nstrips * p.M;
        gidx = gidx < total ? gidx : total - 1;
        strip = gidx / p.M;
        k = (int)(gidx - strip * p.M);
        r = (int)(strip - ((long long)blockIdx.x * p.nconsumer) / p.M);
    }
    strip = strip < last_strip ? strip : last_strip;

    MarchGeom g;
    g.a = p.a0 + strip * p.ls;
    g.n = p.n;
    g.nblk = p.nblk;
    g.stats = p.jpart != nullptr;
    g.node_shift = p.node_shift;

    StreamIO<MC, MODE >= 2, EMIT> io;
    io.sdata = sdata + r * p.run_stride + k;
    io.stage_stride = stage_stride;
    io.M = p.M;
    io.nstages = p.nstages;
    io.full_s = smem_u32(full);
    // output cursor of the first block: site a - 16 - 3 (b = -2), or a - 24 - 11 (b = -3, two steps per pass)
    io.yrow = p.Y + (g.a + FuseGeom<KF>::kOutOff) * p.M + k;   // block b = -2
    io.prow = EMIT ? p.P0out + (g.a + FuseGeom<KF>::kOutOff) : nullptr;
    io.col0 = k == 0;
    io.jp = p.jpart != nullptr ? p.jpart + 2 * k : nullptr;
    io.jstride = p.jstride;
    io.levels = levels;
    io.nnode = p.nnode;
    io.nsm = nsm + tid;
    io.pls_s = smem_u32(pls + tid); io.plc_s = io.pls_s;
    io.nthr = nwarps * 32;
    io.lane = lane;
    io.prod = p.dedicated ? nullptr : &prod;
    io.warp = warp; io.nwarps = nwarps; io.turn = 0;
    io.st = 0;
    io.par = 0;
    io.ready = false; io.mine = false;
    io.stg = 0; io.ahead = p.prefetch; io.nit = prod.nit;
    io.duty = warp == 0 ? nwarps : warp;              // stage 0's duty was done by the prologue
    if (io.duty + p.prefetch >= prod.nit) io.duty = 0x7fffffff;
    io.cur = io.sdata;
    io.stage_bytes = (uint32_t)stage_stride * 8u;
    io.ring_bytes = io.stage_bytes * (uint32_t)p.nstages;
    io.half_bytes = (uint32_t)(8 * p.M) * 8u;
    io.base_s = smem_u32(io.sdata);
    io.nxt_s = io.base_s + io.half_bytes; io.cur_s = io.nxt_s;      // the first block is the second half of stage 0
    io.half = 1; io.need_wait = true;
    io.bar_s = io.full_s; io.bar_end = io.full_s + 8u * (uint32_t)p.nstages;
    io.prev_empty = smem_u32(sink);

    double F = (k == p.M - 1) ? p.Flast : p.F0;
    if (p.Fpp != nullptr) {
        const double2 f = p.Fpp[strip / p.spp];
        F = (k == p.M - 1) ? f.y : f.x;
    }
    if (MODE == 1) l96_march_t<MP, STATS, EulerPipe<MP>>(io, g, F, p.c);
    else if (MODE == 2) l96_march2_t<MP, Rk4Pipe<MP>, STATS>(io, g, F, p.c);
    else if (MODE == 3) l96_march2_t<MP, EulerPipe<MP>, STATS>(io, g, F, p.c);
    else if (MODE == 4) l96_marchk_t<MP, EulerPipe<MP>, 4, STATS>(io, g, F, p.c);
    else l96_march_t<MP, STATS, Rk4Pipe<MP>>(io, g, F, p.c);
    if (p.cycles != nullptr && tid == 0) {
        atomicAdd(p.cycles, (unsigned long long)(clock64() - clk0));
        atomicAdd(p.cycles + 1, 1ULL);
    }
}

// 16 warps per CTA (4 per scheduler, so each thread can hold up to 128 registers), all of
// them consumers: R strips x M columns <= 512 threads.
static int producer_dedicated() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDS_PRODUCER"); v = (e && e[0] == 'i') ? 0 : 1; }
    return v;
}
// Blocks the shared producer keeps in flight.  A block of the two-step kernel takes ~1.4 us, far above the HBM
// latency, so a short distance is enough -- and every stage NOT in flight is slack between the fastest and
// the slowest warp of the CTA (the producer of block c + d reuses the stage of block c + d - nstages, which
// every warp must have left): d = 6 of 8 stages made the fetching warp wait ~1000 clocks for a straggler.
static int stream_prefetch() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FDS_PREFETCH"); v = e ? atoi(e) : 4; if (v < 1) v = 1; }
    return v;
}
int two_step_runs_per_cta(int M) { return M <= 256 ? 256 / M : 0; }
int stream_runs_per_cta(int M) {
    const int maxc = producer_dedicated() ? 480 : 512;
    return M <= maxc ? maxc / M : 0;
}

cudaError_t launch_l96_rk4_stream(const StepArgs& a, const StripPlan& plan, cudaStream_t st, LaunchStats* ls,
                                  int euler) {
    const int nfuse = a.nfuse > 1 ? a.nfuse : ((a.jpart2 != nullptr || a.two_steps) ? 2 : 1);
    if (nfuse != 1 && nfuse != 2 && !(nfuse == 4 && euler)) return cudaErrorInvalidValue;
    const int fuse2 = nfuse >= 2;
    StreamParams p;
    p.X = a.X; p.Y = a.Y; p.jpart = a.jpart; p.nnode = plan.nnode;   // a.jpart: [nnode][M][2]
    p.n = a.n;
    p.node_shift = plan.node_shift;
    p.M = a.M;
    p.R = stream_runs_per_cta(a.M);
    if (p.R < 1) return cudaErrorInvalidValue;            // M too large for this kernel
    p.dedicated = producer_dedicated();
    size_t smem_budget = 226 * 1024;
    if (a.runs_per_cta > 0) {
        if (a.runs_per_cta > p.R) return cudaErrorInvalidValue;
        p.R = a.runs_per_cta;
        if (a.ctas_per_sm >= 2) {
            if (p.R * a.M > 256) return cudaErrorInvalidValue;
            p.dedicated = 0;
            smem_budget = 112 * 1024;
        }
    }
    p.pack = 0;
    if (fuse2) {
        if (p.R * a.M > 256) return cudaErrorInvalidValue;   // set runs_per_cta = two_step_runs_per_cta(M)
        p.dedicated = 0;                                      // all 8 warps compute and take turns issuing the copies
    }
    p.nconsumer = p.R * a.M;
    if (fuse2 && a.M >= 9 && a.M <= 256) {
        // 256 / M strips leave lanes idle (238 of 256 at M = 34): pack the (strip, column) pairs over all 256
        // lanes of every CTA instead; a CTA then touches up to 256 / M + 2 strips and stages all of them
        p.pack = 1;
        p.nconsumer = 256;
        p.R = (256 + a.M - 1) / a.M + 1;
    }
    p.chunk = fuse2 ? 16 : 8;
    int pad = (16 - (7 * a.M) % 16) % 16;
    if (pad & 1) pad += 1;
    p.run_stride = p.chunk * a.M + pad;
    const int cthreads = ((p.nconsumer + 31) / 32) * 32;
    const int block = cthreads + 32 * p.dedicated;
    const size_t stage_bytes = (size_t)p.R * p.run_stride * sizeof(double);
    if (fuse2 && block > 256) return cudaErrorInvalidValue;
    const size_t ns_bytes = (size_t)(2 * nfuse) * (plan.node_shift - kLeafShift) * cthreads * sizeof(double) +
                            (fuse2 ? (size_t)(8 * nfuse) * 256 * sizeof(double) : 0);
    const size_t budget = smem_budget - kStreamHdrBytes - ns_bytes;
    int nst = (int)std::min<size_t>(kMaxStages, budget / stage_bytes);
    if (nst < 2) return cudaErrorInvalidValue;
    p.nstages = nst;
    p.a0 = plan.a0; p.ls = plan.ls; p.nstrips = plan.nstrips; p.nblk = plan.nblk;
    p.F0 = a.F0; p.Flast = a.M > 1 ? a.Flast : a.F0;
    p.Fpp = a.Fpp; p.spp = a.spp > 0 ? a.spp : 1;
    p.jpart2 = a.jpart2; p.fuse2 = fuse2;
    p.P0out = a.primal_out;
    p.cycles = a.cycles;
    const bool emit = a.primal_out != nullptr;
    if (emit && (euler || nfuse != 2 || a.jpart == nullptr)) return cudaErrorInvalidValue;   // RK4, two steps, with objective sums
    p.nfuse = nfuse; p.jstride = (long long)plan.nnode * a.M * 2;
    // (FDS_PREFETCH counts blocks; a stage of the two-block ring is two of them)
    // two-block stages: the copies of stage c + d are issued in the MIDDLE of stage c into the slot of stage c + d - nst,
    // which every warp left when it began stage c + d - nst + 1: d = nst - 1 still leaves half a stage of slack
    // (forward Euler, four steps per pass: three stages fit next to the pair-sum and node tables)
    p.prefetch = fuse2 ? std::max(1, std::min(nst - 1, (stream_prefetch() + 1) / 2))
                       : std::max(1, std::min(nst - 2, stream_prefetch()));
    p.c = Rk4Coef{0.5 * a.dt, a.dt, a.dt / 6.0};
    const int grid = p.pack ? (int)((plan.nstrips * a.M + p.nconsumer - 1) / p.nconsumer)
                            : (int)((plan.nstrips + p.R - 1) / p.R);
    const size_t smem = kStreamHdrBytes + ns_bytes + stage_bytes * nst;
    cudaError_t e = cudaSuccess;
#define FDS_STREAM_LAUNCH2(MP, MC, ST, EU)                                                                         \
    do {                                                                                                           \
        e = cudaFuncSetAttribute(l96_rk4_stream_kernel<MP, MC, ST, EU>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                                       \
        if (e != cudaSuccess) return e;                                                                            \
        l96_rk4_stream_kernel<MP, MC, ST, EU><<<grid, block, smem, st>>>(p);                                       \
    } while (0)
#define FDS_STREAM_LAUNCH2E(MP, MC)                                                                               \
    do {                                                                                                           \
        e = cudaFuncSetAttribute(l96_rk4_stream_kernel<MP, MC, true, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem);                                                                       \
        if (e != cudaSuccess) return e;                                                                            \
        l96_rk4_stream_kernel<MP, MC, true, 2, true><<<grid, block, smem, st>>>(p);                                \
    } while (0)
#define FDS_STREAM_LAUNCH(MP, MC, ST) \
    do { if (emit) { if (ST) FDS_STREAM_LAUNCH2E(MP, MC); } \
         else if (euler && nfuse == 4) FDS_STREAM_LAUNCH2(MP, MC, ST, 4); else if (euler && fuse2) FDS_STREAM_LAUNCH2(MP, MC, ST, 3); \
         else if (euler) FDS_STREAM_LAUNCH2(MP, MC, ST, 1); \
         else if (fuse2) FDS_STREAM_LAUNCH2(MP, MC, ST, 2); \
         else FDS_STREAM_LAUNCH2(MP, MC, ST, 0); } while (0)
#define FDS_STREAM_MC(MP, ST)                                      \
    do {                                                           \
        if (a.M == 34) FDS_STREAM_LAUNCH(MP, 34, ST);              \
        else if (a.M == 18) FDS_STREAM_LAUNCH(MP, 18, ST);         \
        else FDS_STREAM_LAUNCH(MP, 0, ST);                         \
    } while (0)
    const bool stats = p.jpart != nullptr;
#ifdef FDS_DEV_ONE
    // development build (tools/sass_hot.py): only the headline instantiation, seconds instead of a minute to compile
    (void)stats;
    if (a.exact) FDS_STREAM_LAUNCH2(ExactMath, 34, true, 2); else FDS_STREAM_LAUNCH2(FastMath, 34, true, 2);
#else
    if (a.exact) { if (stats) FDS_STREAM_MC(ExactMath, true); else FDS_STREAM_MC(ExactMath, false); }
    else         { if (stats) FDS_STREAM_MC(FastMath, true);  else FDS_STREAM_MC(FastMath, false); }
#endif
#undef FDS_STREAM_MC
#undef FDS_STREAM_LAUNCH
#undef FDS_STREAM_LAUNCH2
#undef FDS_STREAM_LAUNCH2E
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// ------------------------------- simple kernel ------------------------------- //
template <class MP, int EULER>
__global__ void __launch_bounds__(256) l96_simple_kernel(const double* __restrict__ X, double* __restrict__ Y,
                                                         int M, long long lo, long long count, double F0,
                                                         double Flast, Rk4Coef c, const double2* __restrict__ Fpp,
                                                         long long slot) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const long long i = lo + idx / M;
    const int k = (int)(idx % M);
    double F = (k == M - 1) ? Flast : F0;
    if (Fpp != nullptr) {                      // batch: site i belongs to problem i / slot (lo >= 0)
        const double2 f = Fpp[i / slot];
        F = (M > 1 && k == M - 1) ? f.y : f.x;
    }
    if (EULER) {
        double x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = X[(i - 2 + j) * M + k];
        Y[i * M + k] = l96_euler_point<MP>(x, F, c.h);
    } else {
        double x[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) x[j] = X[(i - 8 + j) * M + k];
        Y[i * M + k] = l96_rk4_point<MP>(x, F, c);
    }
}

cudaError_t launch_l96_simple(const StepArgs& a, int euler, cudaStream_t st, LaunchStats* ls) {
    const long long count = (a.hi - a.lo) * a.M;
    if (count <= 0) return cudaSuccess;
    const int block = 256;
    const long long grid = (count + block - 1) / block;
    Rk4Coef c{0.5 * a.dt, a.dt, a.dt / 6.0};
    const double Fl = a.M > 1 ? a.Flast : a.F0;
#define FDS_LAUNCH_SIMPLE(MP, E) \
    l96_simple_kernel<MP, E><<<(unsigned)grid, block, 0, st>>>(a.X, a.Y, a.M, a.lo, count, a.F0, Fl, c, a.Fpp, a.slot)
    if (a.exact) { if (euler) FDS_LAUNCH_SIMPLE(ExactMath, 1); else FDS_LAUNCH_SIMPLE(ExactMath, 0); }
    else         { if (euler) FDS_LAUNCH_SIMPLE(FastMath, 1);  else FDS_LAUNCH_SIMPLE(FastMath, 0); }
#undef FDS_LAUNCH_SIMPLE
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// per-step time derivative of a single-column state, d_i = dt * f(x)_i (analytic time dilation)
template <class MP, int EULER>
__global__ void __launch_bounds__(256) l96_ddt_kernel(const double* __restrict__ X, double* __restrict__ d, long long n,
                                                      double F0, double dt, const double2* __restrict__ Fpp,
                                                      long long slot) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double F = Fpp != nullptr ? Fpp[i / slot].x : F0;
    const double k = EULER ? l96_rhs_fun3d<MP>(X[i - 2], X[i - 1], X[i], X[i + 1], F)
                           : l96_rhs<MP>(X[i - 2], X[i - 1], X[i], X[i + 1], F);
    d[i] = MP::mul(dt, k);
}

cudaError_t launch_l96_ddt(const double* X, double* d, long long n, double F0, double dt, int exact, int euler,
                           const double2* Fpp, long long slot, cudaStream_t st, LaunchStats* ls) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (exact) { if (euler) l96_ddt_kernel<ExactMath, 1><<<grid, 256, 0, st>>>(X, d, n, F0, dt, Fpp, slot);
                 else l96_ddt_kernel<ExactMath, 0><<<grid, 256, 0, st>>>(X, d, n, F0, dt, Fpp, slot); }
    else       { if (euler) l96_ddt_kernel<FastMath, 1><<<grid, 256, 0, st>>>(X, d, n, F0, dt, Fpp, slot);
                 else l96_ddt_kernel<FastMath, 0><<<grid, 256, 0, st>>>(X, d, n, F0, dt, Fpp, slot); }
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// ------------------------------ objective sums ------------------------------- //
__global__ void __launch_bounds__(256) stats_leaf_kernel(const double* __restrict__ Y, int M, long long n,
                                                         double* __restrict__ jpart, long long nleaf) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nleaf * M) return;
    const long long leaf = idx / M;
    const int k = (int)(idx % M);
    TreeAcc<kLeafShift> a1, a2;
    a1.reset(); a2.reset();
    const long long i0 = leaf * kLeafSites;
    for (int j = 0; j < kLeafSites; ++j) {
        const long long i = i0 + j;
        const double v = i < n ? Y[i * M + k] : 0.0;
        a1.push(v);
        a2.push(__dmul_rn(v, v));
    }
    reinterpret_cast<double2*>(jpart)[idx] = make_double2(a1.result(kLeafShift), a2.result(kLeafShift));
}

cudaError_t launch_stats_leaf(const double* Y, int M, long long n, double* jpart, long long nleaf,
                              cudaStream_t st, LaunchStats* ls) {
    const long long total = nleaf * M;
    stats_leaf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Y, M, n, jpart, nleaf);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// One level of the canonical tree: out[g][c] = tree(in[g*LG .. g*LG+LG-1][c]), zero padded.
constexpr int kLG = 256;

// PER consecutive rows of one column folded in the canonical order (adjacent pairs, left operand = lower
// index), all loads issued before the first add: PER independent requests in flight per thread.  The
// binary-counter form below (TreeAcc) gives the same tree but one dependent load at a time from an
// accumulator in local memory -- 1.5 TB/s on the 3.6 GB of node partials of a segment instead of HBM speed.
template <int PER>
__device__ __forceinline__ double tree_rows(const double* __restrict__ in, long long e0, long long n_in, int C, int c) {
    double v[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) v[j] = (e0 + j < n_in) ? in[(e0 + j) * C + c] : 0.0;
#pragma unroll
    for (int s = 1; s < PER; s <<= 1) {
#pragma unroll
        for (int j = 0; j < PER; j += 2 * s) v[j] = v[j] + v[j + s];
    }
    return v[0];
}
__global__ void __launch_bounds__(512, 2) tree_level_kernel(const double* __restrict__ in, long long n_in, int C,
                                                          int Q, double* __restrict__ out) {
    extern __shared__ double sm[];                 // [Q][C]
    const int c = threadIdx.x % C, q = threadIdx.x / C;
    const long long g = blockIdx.x;
    in += (size_t)blockIdx.y * n_in * C;             // batch element
    out += (size_t)blockIdx.y * gridDim.x * C;
    if (q < Q) {
        const int per = kLG / Q;                   // power of two
        int L = 0;
        while ((1 << L) < per) ++L;
        const long long e0 = g * kLG + (long long)q * per;
        double r;
        // 16 rows at a time (the kernel is compiled for up to 512 threads, two CTAs per SM = 64 registers per thread)
        if (per == 32) r = tree_rows<16>(in, e0, n_in, C, c) + tree_rows<16>(in, e0 + 16, n_in, C, c);
        else if (per == 16) r = tree_rows<16>(in, e0, n_in, C, c);
        else if (per == 64)
            r = (tree_rows<16>(in, e0, n_in, C, c) + tree_rows<16>(in, e0 + 16, n_in, C, c)) +
                (tree_rows<16>(in, e0 + 32, n_in, C, c) + tree_rows<16>(in, e0 + 48, n_in, C, c));
        else if (per == 8) r = tree_rows<8>(in, e0, n_in, C, c);
        else {
            TreeAcc<8> acc;
            acc.reset();
            for (int j = 0; j < per; ++j) {
                const long long e = e0 + j;
                acc.push(e < n_in ? in[e * C + c] : 0.0);
            }
            r = acc.result(L);
        }
        sm[q * C + c] = r;
    }
    __syncthreads();
    for (int s = 1; s < Q; s <<= 1) {
        if (q < Q && (q % (2 * s)) == 0) sm[q * C + c] = sm[q * C + c] + sm[(q + s) * C + c];
        __syncthreads();
    }
    if (q == 0) out[g * C + c] = sm[c];
}

cudaError_t launch_tree_reduce(const double* in, long long n_in, int C, double* out, double* scratch,
                               cudaStream_t st, LaunchStats* ls) {
    return launch_tree_reduce_batched(in, n_in, C, 1, out, scratch, st, ls);
}

// batch of independent reductions: in[batch][n_in][C] -> out[batch][C]
// scratch holds 2 * batch * ceil(n_in/256) * C doubles
cudaError_t launch_tree_reduce_batched(const double* in, long long n_in, int C, long long batch, double* out,
                                       double* scratch, cudaStream_t st, LaunchStats* ls) {
    int Q = 1;
    // <= 512 threads at <= 64 registers: three CTAs of 272 threads per SM at C = 68 (one CTA of 544 before: the
    // pass ran at 4.5 TB/s with 17 warps of loads in flight)
    while (Q * 2 * C <= 512 && Q * 2 <= kLG) Q *= 2;
    if (C > 512) return cudaErrorInvalidValue;
    const long long cap = (n_in + kLG - 1) / kLG;   // size of one scratch half, in rows
    const double* src = in;
    long long n = n_in;
    int flip = 0;
    if (batch < 1) return cudaSuccess;
    for (;;) {
        const long long n_out = (n + kLG - 1) / kLG;
        double* dst = (n_out == 1) ? out : scratch + (size_t)flip * cap * C * batch;
        dim3 grid((unsigned)n_out, (unsigned)batch);
        tree_level_kernel<<<grid, Q * C, (size_t)Q * C * sizeof(double), st>>>(src, n, C, Q, dst);
        if (ls) ls->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_out == 1) break;
        src = dst; n = n_out; flip ^= 1;
    }
    return cudaSuccess;
}

// cross-rank part of the canonical tree (see kernels.h)
__global__ void __launch_bounds__(256) rank_tree_sum_kernel(const double* __restrict__ g, int world, int P,
                                                            long long len, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double v[64];
    for (int r = 0; r < P; ++r) v[r] = r < world ? g[(size_t)r * len + e] : 0.0;
    for (int s = 1; s < P; s <<= 1)
        for (int j = 0; j < P; j += 2 * s) v[j] = __dadd_rn(v[j], v[j + s]);
    out[e] = v[0];
}

cudaError_t launch_rank_tree_sum(const double* gathered, int world, long long len, double* out, cudaStream_t st,
                                 LaunchStats* ls) {
    int P = 1;
    while (P < world) P *= 2;
    if (P > 64 || len < 1) return cudaErrorInvalidValue;
    rank_tree_sum_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(gathered, world, P, len, out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// --------------------------------- halo wrap --------------------------------- //
__global__ void __launch_bounds__(256) halo_wrap_kernel(double* __restrict__ A, int M, long long n,
                                                        long long gl, long long gr) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (gl + gr) * M;
    if (idx >= total) return;
    const long long s = idx / M;
    const int k = (int)(idx % M);
    const long long g = s < gl ? s - gl : n + (s - gl);     // ghost site index
    long long src = g % n;
    if (src < 0) src += n;
    A[g * M + k] = A[src * M + k];
}

cudaError_t launch_halo_wrap(double* A, int M, long long n, long long gl, long long gr, cudaStream_t st,
                             LaunchStats* ls) {
    const long long total = (gl + gr) * M;
    if (total <= 0) return cudaSuccess;
    halo_wrap_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, M, n, gl, gr);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) halo_wrap_batch_kernel(double* __restrict__ A, int M, long long n,
                                                              long long gl, long long gr, long long slot) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (gl + gr) * M;
    if (idx >= total) return;
    A += (size_t)blockIdx.y * slot * M;
    const long long s = idx / M;
    const int k = (int)(idx % M);
    const long long g = s < gl ? s - gl : n + (s - gl);
    long long src = g % n;
    if (src < 0) src += n;
    A[g * M + k] = A[src * M + k];
}

cudaError_t launch_halo_wrap_batch(double* A, int M, long long n, long long gl, long long gr, int batch,
                                   long long slot, cudaStream_t st, LaunchStats* ls) {
    const long long total = (gl + gr) * M;
    if (total <= 0 || batch < 1) return cudaSuccess;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)batch);
    halo_wrap_batch_kernel<<<grid, 256, 0, st>>>(A, M, n, gl, gr, slot);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// one thread per (step, problem, column-sum): canonical tree over the problem's leaves
__global__ void __launch_bounds__(128) batch_jreduce_kernel(const double* __restrict__ leaves, long long nleaf_total,
                                                            long long leaf0, long long slot_leaves, long long leaves_pp,
                                                            int C, int batch, long long steps, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= steps * batch * C) return;
    const int c = (int)(idx % C);
    const long long sp = idx / C;
    const int p = (int)(sp % batch);
    const long long step = sp / batch;
    const double* in = leaves + ((size_t)step * nleaf_total + leaf0 + (size_t)p * slot_leaves) * C + c;
    int L = 0;
    while ((1LL << L) < leaves_pp) ++L;
    TreeAcc<24> acc;
    acc.reset();
    if (L >= 4) {
        // 16 leaves at a time, all loads issued before the first add (tree_rows); the binary counter then only
        // combines the 16-leaf subtrees -- the same canonical tree, without one dependent load per leaf
        for (long long j = 0; j < (1LL << L); j += 16) acc.push(tree_rows<16>(in, j, leaves_pp, C, 0));
        out[idx] = acc.result(L - 4);
    } else {
        for (long long j = 0; j < (1LL << L); ++j) acc.push(j < leaves_pp ? in[(size_t)j * C] : 0.0);
        out[idx] = acc.result(L);
    }
}

cudaError_t launch_batch_jreduce(const double* leaves, long long nleaf_total, long long leaf0, long long slot_leaves,
                                 long long leaves_pp, int C, int batch, long long steps, double* out,
                                 cudaStream_t st, LaunchStats* ls) {
    const long long total = steps * batch * C;
    if (total <= 0) return cudaSuccess;
    batch_jreduce_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(leaves, nleaf_total, leaf0, slot_leaves,
                                                                          leaves_pp, C, batch, steps, out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// Batch: finite-difference sensitivities of the objective on the device, so that only KBs per
// problem leave the GPU.  jsum[steps][nb][M][2] (sums over sites) ->
//   J0[p][t][q]  = jsum[t][p][0][q] / n                                    (primal objective history)
//   G[p][j][q]   = trapez_mean_t(J[t][p][1+j][q] - J0[p][t][q]) / eps,  j = 0..m (row m: inhomogeneous)
// with the reference's trapez_mean (fds/segment.py:9-12): (sum_t d_t + sum_{t<n-1} d_t + (2 d_0 - d_1)) / (2n),
// sums accumulated in step order like numpy's reduction over the leading axis.
__global__ void __launch_bounds__(128) batch_sens_kernel(const double* __restrict__ jsum, int nb, int M,
                                                         long long steps, double n_sites, double eps,
                                                         double* __restrict__ J0, double* __restrict__ G) {
    const int p = blockIdx.x;
    const int MO = M - 1;
    for (int e = threadIdx.x; e < MO * 2; e += blockDim.x) {
        const int j = e >> 1, q = e & 1;
        double sum_all = 0.0, sum_head = 0.0, d0 = 0.0, d1 = 0.0;
        for (long long t = 0; t < steps; ++t) {
            const double* row = jsum + (((size_t)t * nb + p) * M) * 2;
            const double j0 = row[q] / n_sites;
            const double d = __dsub_rn(row[(1 + j) * 2 + q] / n_sites, j0);
            if (t == 0) d0 = d;
            if (t == 1) d1 = d;
            sum_all = t == 0 ? d : __dadd_rn(sum_all, d);
            if (t < steps - 1) sum_head = t == 0 ? d : __dadd_rn(sum_head, d);
        }
        const double ghost = __dsub_rn(__dmul_rn(2.0, d0), d1);
        const double tm = __dadd_rn(__dadd_rn(sum_all, sum_head), ghost) / (double)(2 * steps);
        G[((size_t)p * MO + j) * 2 + q] = tm / eps;
    }
    for (long long e = threadIdx.x; e < steps * 2; e += blockDim.x) {
        const long long t = e >> 1;
        const int q = (int)(e & 1);
        J0[((size_t)p * steps + t) * 2 + q] = jsum[(((size_t)t * nb + p) * M) * 2 + q] / n_sites;
    }
}

cudaError_t launch_batch_sens(const double* jsum, int nb, int M, long long steps, double n_sites, double eps,
                              double* J0, double* G, cudaStream_t st, LaunchStats* ls) {
    if (steps < 2) return cudaErrorInvalidValue;
    batch_sens_kernel<<<nb, 128, 0, st>>>(jsum, nb, M, steps, n_sites, eps, J0, G);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

}  // namespace fds
