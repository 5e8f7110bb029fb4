// K time steps per pass (K >= 2): the two-step march of l96_march2.cuh with a chain of K pipes.
//
// Pipe j + 1 is fed the value that left pipe j ONE PUSH EARLIER, so the stage computations of one push are all
// independent and every step lags its predecessor by exactly 8 sites: for a push of input site i the step-(j+1)
// value of site i - 7 - 8j comes out.  A site is read once and written once per K time steps (16 / K bytes of
// HBM traffic per trajectory-site-step), all K steps' objective sums are accumulated on the way in the specified
// tree, and the result is bit-identical to K single-step passes (tests/test_emulation.py, GPU parity tests).
//
// Used for forward Euler with K = 4 (the reference's own Lorenz-96 solver, tests/solvers/mock_fun3d/fun3d:48-57:
// 10 doubles of pipeline state and 10.5 FP64 instructions per update, so four chained pipes fit the register
// file and the pass stays on the memory roof).  The RK4 kernel keeps its hand-tuned two-step body.
//
// Block b of a strip that starts at site a consumes the input sites a + 8b + 4 .. a + 8b + 11 and produces the
// step-(j+1) values of sites a + 8(b-j) - 3 .. a + 8(b-j) + 4; only step K is stored.  Blocks run from
// b = -(K+1) (warm-up) to b = nblk + K - 1 (drain).  IO policy: as in l96_march2.cuh.
#pragma once
#include "l96_march2.cuh"

namespace fds {

template <int K>
struct FuseGeom {
    static constexpr int kFirst = -(K + 1);            // first block index
    static constexpr int kExtra = 2 * K + 1;           // blocks per strip = nblk + kExtra
    static constexpr int kInOff = 8 * kFirst + 4;      // first input site, relative to the strip start
    static constexpr int kOutOff = -16 * K - 3;        // output cursor of the first block, relative to the strip start
    static constexpr int kPad = 12 + 8 * (K - 1);      // sites the arrays extend beyond the strips on either side
};
// the one-step march (l96_march.cuh: b = -2 .. nblk) and the two-step march (l96_march2.cuh: b = -3 .. nblk + 1)
static_assert(FuseGeom<1>::kFirst == -2 && FuseGeom<1>::kExtra == 3 && FuseGeom<1>::kInOff == -12 &&
              FuseGeom<1>::kOutOff == -19, "one step per pass");
static_assert(FuseGeom<2>::kFirst == kFuse2First && FuseGeom<2>::kInOff == -20 && FuseGeom<2>::kOutOff == -35 &&
              FuseGeom<2>::kPad == 20, "two steps per pass");
static_assert(FuseGeom<kMaxFuse>::kPad + 8 == kPlanPad, "the strip plan pads for the deepest pass");

template <class MP, class PIPE, int K, int P, bool FAST, bool STATS, class IO>
FDS_HD void marchk_site(IO& io, PIPE (&p)[K], double (&xprev)[K], StepStats (&st)[K], long long jb, long long slo,
                        long long shi, long long jlo, long long jhi, double F, const Rk4Coef& c) {
    constexpr int Q = (P + 5) & 7;       // output site of every step, mod 8 (each step lags the previous one by 8)
    double out[K];
    // the later pipes first: they take what the earlier ones produced in the PREVIOUS push
#pragma unroll
    for (int j = K - 1; j >= 1; --j) out[j] = p[j].template push<P>(xprev[j - 1], F, c);
    out[0] = p[0].template push<P>(io.load(P), F, c);
#pragma unroll
    for (int j = 0; j + 1 < K; ++j) xprev[j] = out[j];
    if (FAST) {
        io.store(P, out[K - 1]);
    } else {
        const long long s = jb - 8LL * (K - 1) + P;
        if (s >= slo && s < shi) io.store(P, out[K - 1]);
    }
    if (STATS) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double v = out[j];
            if (!FAST) {
                const long long s = jb - 8LL * j + P;
                v = (s >= jlo && s < jhi) ? v : 0.0;
            }
            const double r1 = st[j].t1.template push<Q>(v);
            const double r2 = st[j].t2.template push<Q>(MP::mul(v, v));
            if (Q == 7) { st[j].b1 = r1; st[j].b2 = r2; }
        }
    }
}

// pair sums of steps J + 1, J, .. 1 (compile-time recursion: the step selects the storage of its pair sums)
template <int BP, int J, class IO, int K>
FDS_HD void marchk_pair(IO& io, StepStats (&st)[K], int b) {
    march2_pair<(BP >= 0 ? BP ^ ((J + 1) & 1) : -1), J>(io, st[J], (b - 1 - J) & 7);
    if constexpr (J > 0) marchk_pair<BP, J - 1>(io, st, b);
}

// BP: parity of b when known at compile time (unchecked blocks only), -1 otherwise
template <class MP, class PIPE, int K, bool FAST, bool STATS, int BP = -1, class IO>
FDS_HD void marchk_block(IO& io, PIPE (&p)[K], double (&xprev)[K], StepStats (&st)[K], const MarchGeom& g, int b,
                         long long slo, long long shi, long long jlo, long long jhi, double F, const Rk4Coef& c) {
    const long long jb = g.a + 8LL * b - 3;      // first step-1 site of this block
#define FDS_MK(P) marchk_site<MP, PIPE, K, P, FAST, STATS>(io, p, xprev, st, jb, slo, shi, jlo, jhi, F, c)
    FDS_MK(0);
    FDS_MK(1);
    FDS_MK(2);
    // step j + 1: the 8-block b - 1 - j of the strip is complete (Q == 7 at P == 2)
    if (FAST) {
        if (STATS) {
            // 8-block b - 1 - j has parity BP ^ ((1 + j) & 1): an even block parks its sum, an odd one pairs it
            // and stores the pair (march2_pair)
            marchk_pair<BP, K - 1>(io, st, b);
        }
    } else if (STATS) {
#pragma unroll
        for (int j = K - 1; j >= 0; --j)
            if (b - 1 - j >= 0 && b - 1 - j < g.nblk) march2_leaf(io, st[j], g, j, b - 1 - j);
    }
    FDS_MK(3);
    FDS_MK(4);
    io.template mid_block<BP>();
    FDS_MK(5);
    FDS_MK(6);
    FDS_MK(7);
#undef FDS_MK
    if (FAST && STATS) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < K; ++j) any = any || ((b - 1 - j) & 7) == 7;
        if (any) {
#pragma unroll
            for (int j = K - 1; j >= 0; --j)
                if (((b - 1 - j) & 7) == 7) march2_node(io, st[j], g, j, b - 1 - j);
        }
    }
}

template <class MP, class PIPE, int K, bool STATS, class IO>
FDS_HD void l96_marchk_t(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    static_assert(PIPE::kLag == 3, "the chaining assumes a lag of 3 sites per pipe");
    static_assert(K >= 2 && K <= kMaxFuse, "steps per pass");
    PIPE p[K];
    StepStats st[K];
    double xprev[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { p[j].reset(); st[j].reset(); xprev[j] = 0.0; }
    const long long slo = g.a, shi = g.a + 8LL * g.nblk;
    const long long jlo = slo > 0 ? slo : 0;
    const long long jhi = g.n < shi ? g.n : shi;
    const long long nfl = g.n & ~(long long)(kLeafSites - 1);
    const long long ncl = (g.n + kLeafSites - 1) & ~(long long)(kLeafSites - 1);
    // Unchecked blocks [bfl, bfh]: the 8 step-K outputs (sites a + 8(b-K) + 5 ..) are stored by this strip, every
    // step's leaf exists (0 <= b - 1 - j < nblk), and no step's 8 values lie in the partial leaf containing site n.
    int bfl = K, bfh = g.nblk;
    if (nfl != ncl) {
        // clean: the highest site (step 1: a + 8b + 4) below nfl, or the lowest (step K: a + 8b - 3 - 8(K-1)) at or above ncl
        const long long below = floor_div(nfl - g.a - 5, 8);
        const long long above = -floor_div(-(ncl - g.a + 3 + 8 * (K - 1)), 8);
        if (below >= bfl && below < bfh) {
            if (above > bfh || (below - bfl) >= (bfh - above)) bfh = (int)below; else bfl = (int)above;
        } else if (below < bfl) {
            if (above > bfl) bfl = (int)(above < (long long)bfh + 1 ? above : (long long)bfh + 1);
        }
    }
    bfl = io.warp_max(bfl);
    bfh = io.warp_min(bfh);
    bfl += bfl & 1;                 // the unchecked body starts on an even block and handles two blocks per iteration
    const int last = g.nblk + K - 1;
    int b = FuseGeom<K>::kFirst;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int stop = pass == 0 ? (bfl < last + 1 ? bfl : last + 1) : last + 1;
#pragma unroll 1
        for (; b < stop; ++b) {
            io.begin_block(b);
            marchk_block<MP, PIPE, K, false, STATS>(io, p, xprev, st, g, b, slo, shi, jlo, jhi, F, c);
            io.end_block(b);
        }
        if (pass == 0) {
#pragma unroll 1
            for (; b + 1 <= bfh; b += 2) {
                io.template begin_block<0>(b);
                marchk_block<MP, PIPE, K, true, STATS, 0>(io, p, xprev, st, g, b, slo, shi, jlo, jhi, F, c);
                io.template end_block<0>(b);
                io.template begin_block<1>(b + 1);
                marchk_block<MP, PIPE, K, true, STATS, 1>(io, p, xprev, st, g, b + 1, slo, shi, jlo, jhi, F, c);
                io.template end_block<1>(b + 1);
            }
        }
    }
}

// four forward-Euler steps per pass
template <class MP, class IO>
FDS_HD void l96_euler_march4(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_marchk_t<MP, EulerPipe<MP>, 4, true>(io, g, F, c);
    else l96_marchk_t<MP, EulerPipe<MP>, 4, false>(io, g, F, c);
}
// the generic march with K = 2 (checked against l96_march2.cuh in the emulation tests)
template <class MP, class IO>
FDS_HD void l96_rk4_marchk2(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_marchk_t<MP, Rk4Pipe<MP>, 2, true>(io, g, F, c);
    else l96_marchk_t<MP, Rk4Pipe<MP>, 2, false>(io, g, F, c);
}

}  // namespace fds
