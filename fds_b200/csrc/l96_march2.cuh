// Two time steps per pass: the per-thread body of the temporally blocked streaming kernel (RK4, or forward
// Euler through the same push/lag interface).
//
// A second pipe (Rk4Pipe / EulerPipe) is chained behind the first: the step-1 value that leaves pipe 1 (site i-7 for a
// push of site i) is pushed into pipe 2 ONE PUSH LATER (so that the eight stage computations of a push
// -- four per pipe -- are all independent), whose output is the step-2 value at site i-15.  The
// intermediate state never touches shared or global memory, so a site is read once and written once
// per TWO time steps: 8 B of HBM traffic per trajectory-site-step instead of 16.  Both steps'
// objective sums are accumulated on the way (the intermediate state's from pipe 1's output) in the
// same specified tree, so the result is bit-identical to two single-step passes.
//
// Block b of a strip that starts at site a consumes the 8 input sites a + 8b + 4 .. a + 8b + 11,
// produces the step-1 values of sites a + 8b - 3 .. a + 8b + 4 (never stored) and the step-2 values
// of sites a + 8b - 11 .. a + 8b - 4.  Blocks run from b = -3 (warm-up: x'' is exact from site a - 4
// on) to b = nblk + 1 (drain of the 15-site lag).
//
// IO policy: as in l96_march.cuh, with the output cursor at site a + 8b - 11 and the per-step
// variants   emit_node2(step, node, s1, s2),   ns_get2(step, level, which),   ns_put2(step, level, which, v),
// plus   mid_block()  (called once inside EVERY block: ring bookkeeping, probe of the next block's input)   and
// warp_max(int) / warp_min(int)  (reductions over the lanes that share a code path; identity for one thread).
#pragma once
#include "l96_march.cuh"

namespace fds {

constexpr int kFuse2First = -3;                              // first block index
FDS_HD int march2_iterations(int nblk) { return nblk + 5; }  // b = -3 .. nblk + 1

template <class IO>
struct NsStep {            // NodeCounter storage view of one of the two steps
    IO& io;
    int step;
    FDS_HD double ns_get(int l, int q) const { return io.ns_get2(step, l, q); }
    FDS_HD void ns_put(int l, int q, double v) { io.ns_put2(step, l, q, v); }
};

// 64-site leaf of one objective sum, kept as the four pair sums q_j = (block 2j) + (block 2j + 1) of its eight
// 8-site blocks: an even block parks its sum in a register, the odd block adds its own and stores the pair in
// per-thread storage of the IO policy (shared memory in the CUDA kernel); the leaf (q0 + q1) + (q2 + q3) -- the
// balanced tree of oracle/plugins.py tree_sum -- is formed when the eighth block completes.  Compared with a
// binary counter in registers (LeafStack) the hot loop carries no select instructions: one add and one store
// per pair of blocks and sum.
//   void pl_put(int sum, int slot, double v) / double pl_get(int sum, int slot)      sum = 2 * step + which
struct StepStats {         // objective accumulators of one step
    Tree8 t1, t2;
    double e1, e2;         // parked 8-site sums of the even block
    NodeCounter ns;
    double b1, b2;
    FDS_HD void reset() {
        t1.t1 = t1.t2 = t1.t4 = 0.0; t2 = t1;
        e1 = e2 = 0.0;
        ns.reset();
        b1 = b2 = 0.0;
    }
};

// 8-block j (0 .. 7) of a leaf of step STEP is complete (its sums in s.b1, s.b2); PAR = parity of j if known at
// compile time.  CUR: the unchecked two-block body keeps a cursor on the slot of the current pair (the two steps
// pair the SAME 8-block index inside one loop iteration), so the store needs no address arithmetic.
template <int PAR, int STEP, bool CUR = false, class IO>
FDS_HD void march2_pair(IO& io, StepStats& s, int j) {
    if (PAR == 0 || (PAR < 0 && !(j & 1))) { s.e1 = s.b1; s.e2 = s.b2; return; }
    if (CUR) {
        io.template pl_cur_put<2 * STEP>(s.e1 + s.b1);
        io.template pl_cur_put<2 * STEP + 1>(s.e2 + s.b2);
    } else {
        io.pl_put(2 * STEP, j >> 1, s.e1 + s.b1);
        io.pl_put(2 * STEP + 1, j >> 1, s.e2 + s.b2);
    }
}

// the eighth block of a leaf is complete and paired: fold the leaf into the node counter
template <class IO>
FDS_HD void march2_node(IO& io, StepStats& s, const MarchGeom& g, int step, int k) {
    double s1 = (io.pl_get(2 * step, 0) + io.pl_get(2 * step, 1)) + (io.pl_get(2 * step, 2) + io.pl_get(2 * step, 3));
    double s2 = (io.pl_get(2 * step + 1, 0) + io.pl_get(2 * step + 1, 1)) +
                (io.pl_get(2 * step + 1, 2) + io.pl_get(2 * step + 1, 3));
    const long long leaf = (g.a + 8LL * k) >> kLeafShift;
    if (leaf < 0 || leaf * kLeafSites >= g.n) { s1 = 0.0; s2 = 0.0; }     // leaves outside [0, n): exact zeros
    NsStep<IO> view{io, step};
    if (s.ns.push(view, s1, s2, g.node_shift - kLeafShift))
        io.emit_node2(step, leaf >> (g.node_shift - kLeafShift), s1, s2);
}

// 8-block k of the strip (sites a + 8k .. a + 8k + 7) of step `step` is complete: checked path
template <class IO>
FDS_HD void march2_leaf(IO& io, StepStats& s, const MarchGeom& g, int step, int k) {
    // (STEP only selects the storage of the pair sums: dispatch the run-time step of the K-step march)
    switch (step) {
        case 0: march2_pair<-1, 0>(io, s, k & 7); break;
        case 1: march2_pair<-1, 1>(io, s, k & 7); break;
        case 2: march2_pair<-1, 2>(io, s, k & 7); break;
        default: march2_pair<-1, 3>(io, s, k & 7); break;
    }
    if ((k & 7) == 7) march2_node(io, s, g, step, k);
}

// Unchecked blocks: the pair sums advance branch-free right after site 2 (march2_block) and the leaf that completed
// in this block -- one block in eight per step -- is folded into the node counter at the END of the block, so the
// eight pushes form ONE basic block.
template <class MP, class PIPE, int P, bool FAST, bool STATS, class IO>
FDS_HD void march2_site(IO& io, PIPE& p1, PIPE& p2, double& x1_prev, StepStats& sa, StepStats& sb,
                        long long j1, long long j2, long long slo, long long shi, long long jlo, long long jhi,
                        double F, const Rk4Coef& c) {
    // pipe 2 is fed the step-1 value of the PREVIOUS push (site i - 8 for pipe 1's index i = P mod 8):
    // (i - 8) = i2 + 4  =>  i2 = P (mod 8)
    constexpr int Q1 = (P + 5) & 7;      // step-1 output site i - 3, mod 8
    constexpr int Q2 = (P + 5) & 7;      // step-2 output site i - 11, mod 8
    const double x2 = p2.template push<P>(x1_prev, F, c);
    const double x1 = p1.template push<P>(io.load(P), F, c);
    x1_prev = x1;
    if (FAST) {
        io.store(P, x2);
        if (STATS) {
            const double r1 = sa.t1.template push<Q1>(x1);
            const double r2 = sa.t2.template push<Q1>(MP::mul(x1, x1));
            if (Q1 == 7) { sa.b1 = r1; sa.b2 = r2; }
            const double q1 = sb.t1.template push<Q2>(x2);
            const double q2 = sb.t2.template push<Q2>(MP::mul(x2, x2));
            if (Q2 == 7) { sb.b1 = q1; sb.b2 = q2; }
        }
    } else {
        const long long s1 = j1 + P, s2 = j2 + P;
        if (s2 >= slo && s2 < shi) io.store(P, x2);
        if (STATS) {
            const double v1 = (s1 >= jlo && s1 < jhi) ? x1 : 0.0;
            const double r1 = sa.t1.template push<Q1>(v1);
            const double r2 = sa.t2.template push<Q1>(MP::mul(v1, v1));
            if (Q1 == 7) { sa.b1 = r1; sa.b2 = r2; }
            const double v2 = (s2 >= jlo && s2 < jhi) ? x2 : 0.0;
            const double q1 = sb.t1.template push<Q2>(v2);
            const double q2 = sb.t2.template push<Q2>(MP::mul(v2, v2));
            if (Q2 == 7) { sb.b1 = q1; sb.b2 = q2; }
        }
    }
}

// BP: parity of b when known at compile time (unchecked blocks only), -1 otherwise
template <class MP, class PIPE, bool FAST, bool STATS, int BP = -1, class IO>
FDS_HD void march2_block(IO& io, PIPE& p1, PIPE& p2, double& x1_prev, StepStats& sa, StepStats& sb,
                         const MarchGeom& g, int b, long long slo, long long shi, long long jlo, long long jhi,
                         double F, const Rk4Coef& c) {
    const long long j1 = g.a + 8LL * b - 3, j2 = g.a + 8LL * b - 11;
#define FDS_M2(P) march2_site<MP, PIPE, P, FAST, STATS>(io, p1, p2, x1_prev, sa, sb, j1, j2, slo, shi, jlo, jhi, F, c)
    FDS_M2(0);
    FDS_M2(1);
    FDS_M2(2);
    if (FAST) {
        // unchecked blocks have 2 <= b <= nblk: both leaves exist
        if (STATS) {
            // step 2 counts 8-block b - 2 (parity BP), step 1 counts 8-block b - 1 (the other parity)
            march2_pair<BP, 1, (BP >= 0)>(io, sb, (b - 2) & 7);
            march2_pair<(BP >= 0 ? 1 - BP : -1), 0, (BP >= 0)>(io, sa, (b - 1) & 7);
        }
    } else {
        // step 2: sites a + 8(b-2) .. a + 8(b-2) + 7 are complete (Q2 == 7 at P == 2)
        if (STATS && b >= 2 && b - 2 < g.nblk) march2_leaf(io, sb, g, 1, b - 2);
        // step 1: sites a + 8(b-1) .. a + 8(b-1) + 7 are complete (Q1 == 7 at P == 2)
        if (STATS && b >= 1 && b - 1 < g.nblk) march2_leaf(io, sa, g, 0, b - 1);
    }
    FDS_M2(3);
    FDS_M2(4);
    io.template mid_block<BP>();   // ring bookkeeping + early probe of the next stage's input, hidden in the FP64 stream
    FDS_M2(5);
    FDS_M2(6);
    FDS_M2(7);
#undef FDS_M2
    // (unchecked blocks: the leaves that complete are folded into the node counters by the caller, once per PAIR of
    // blocks -- l96_march2_t)
}

// PIPE: Rk4Pipe<MP> or EulerPipe<MP> (any pipe with push x[i+4] -> x'[i-3]: the chaining below only uses the lag)
template <class MP, class PIPE, bool STATS, class IO>
FDS_HD void l96_march2_t(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    static_assert(PIPE::kLag == 3, "the two-step chaining assumes a lag of 3 sites per pipe");
    PIPE p1, p2;
    p1.reset();
    p2.reset();
    StepStats sa, sb;
    sa.reset();
    sb.reset();
    double x1_prev = 0.0;
    const long long slo = g.a, shi = g.a + 8LL * g.nblk;
    const long long jlo = slo > 0 ? slo : 0;
    const long long jhi = g.n < shi ? g.n : shi;
    const long long nfl = g.n & ~(long long)(kLeafSites - 1);
    const long long ncl = (g.n + kLeafSites - 1) & ~(long long)(kLeafSites - 1);
    // Blocks b in [bfl, bfh] take the unchecked path: the 8 step-2 outputs (sites a + 8b - 11 ..) are stored by
    // this strip and neither they nor the 8 step-1 values (a + 8b - 3 ..) lie in the partial leaf that contains
    // site n (leaves wholly outside [0, n) are zeroed as a whole).  Computed once: the loop compares two ints.
    int bfl = 2, bfh = g.nblk;                                  // j2 >= slo  <=>  b >= 2 ;  j2 + 8 <= shi  <=>  b <= nblk
    if (nfl != ncl) {                                           // a partial leaf [nfl, ncl) exists
        // clean blocks: j1 + 8 <= nfl  (b <= (nfl - a - 5) / 8)   or   j2 >= ncl  (b >= (ncl - a + 11) / 8, rounded up)
        const long long below = floor_div(nfl - g.a - 5, 8), above = -floor_div(-(ncl - g.a + 11), 8);
        if (below >= bfl && below < bfh) {
            // the strip straddles the partial leaf: keep the longer clean side unchecked
            if (above > bfh || (below - bfl) >= (bfh - above)) bfh = (int)below; else bfl = (int)above;
        } else if (below < bfl) {
            if (above > bfl) bfl = (int)(above < (long long)bfh + 1 ? above : (long long)bfh + 1);
        }
    }
    // a warp stays on ONE code path: the unchecked range common to all its lanes (they may sit on two strips),
    // computed once, so the loops below compare warp-uniform integers and carry no vote
    bfl = io.warp_max(bfl);
    bfh = io.warp_min(bfh);
    bfl += bfl & 1;                 // the unchecked body starts on an even block (the checked path is valid for every block)
    const int last = g.nblk + 1;
    int b = kFuse2First;
    // checked head [first, bfl), unchecked body [bfl, bfh], checked tail (bfh, last]: pass 0 = head + body, pass 1 = tail
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int stop = pass == 0 ? (bfl < last + 1 ? bfl : last + 1) : last + 1;
#pragma unroll 1
        for (; b < stop; ++b) {
            io.begin_block(b);
            march2_block<MP, PIPE, false, STATS>(io, p1, p2, x1_prev, sa, sb, g, b, slo, shi, jlo, jhi, F, c);
            io.end_block(b);
        }
        if (pass == 0) {
            // Unchecked body, two blocks per iteration so that the lowest bit of the leaf counters' position is a
            // compile-time constant (b is even here; a single last block is left to the checked tail).
            // In iteration b the two steps pair the same 8-block index b - 1 (step 1 in block b, step 2 in block b + 1),
            // so one cursor serves both and BOTH leaves complete in the same iteration (b = 0 mod 8): one rare
            // branch per pair of blocks folds them into the node counters.  (Step 1 parks 8-block b in a register
            // during block b + 1: the stored pairs are untouched until block b + 2.)
            if (STATS) io.pl_cur_set(((b - 1) & 7) >> 1);
#pragma unroll 1
            for (; b + 1 <= bfh; b += 2) {
                // (block parity = which half of its ring stage the block is: the IO policy synchronises once per pair)
                io.template begin_block<0>(b);
                march2_block<MP, PIPE, true, STATS, 0>(io, p1, p2, x1_prev, sa, sb, g, b, slo, shi, jlo, jhi, F, c);
                io.template end_block<0>(b);
                io.template begin_block<1>(b + 1);
                march2_block<MP, PIPE, true, STATS, 1>(io, p1, p2, x1_prev, sa, sb, g, b + 1, slo, shi, jlo, jhi, F, c);
                io.template end_block<1>(b + 1);
                if (STATS) {
                    io.pl_cur_next();
                    if ((b & 7) == 0) {
                        march2_node(io, sa, g, 0, b - 1);
                        march2_node(io, sb, g, 1, b - 1);
                        io.pl_cur_set(0);
                    }
                }
            }
        }
    }
}

template <class MP, bool STATS, class IO>
FDS_HD void l96_rk4_march2_t(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    l96_march2_t<MP, Rk4Pipe<MP>, STATS>(io, g, F, c);
}

template <class MP, class IO>
FDS_HD void l96_rk4_march2(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_march2_t<MP, Rk4Pipe<MP>, true>(io, g, F, c);
    else l96_march2_t<MP, Rk4Pipe<MP>, false>(io, g, F, c);
}

// two forward-Euler steps per pass (the reference's own Lorenz-96 solver, tests/solvers/mock_fun3d/fun3d:48-57)
template <class MP, class IO>
FDS_HD void l96_euler_march2(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_march2_t<MP, EulerPipe<MP>, true>(io, g, F, c);
    else l96_march2_t<MP, EulerPipe<MP>, false>(io, g, F, c);
}

}  // namespace fds
