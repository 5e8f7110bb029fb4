// Lorenz-96 arithmetic shared by the CUDA kernels and the host emulation harness.
//
// Operation order is the contract with oracle/plugins.py (l96_rhs_factored,
// l96_rhs_fun3d, l96_rk4_step, l96_euler_step, tree_sum): in ExactMath mode every
// value below is bit-identical to the numpy oracle.
#pragma once
#include "common.cuh"

namespace fds {

// dx_i = ((x_{i+1} - x_{i-2}) * x_{i-1} - x_i) + F      (RK4 plugin; reference
// tests/solvers/Lorenz96/Lorenz96.f90:20)
template <class MP>
FDS_HD double l96_rhs(double xm2, double xm1, double x0, double xp1, double F) {
    if (MP::kExact) {
        double t = MP::sub(xp1, xm2);
        t = MP::mul(t, xm1);
        t = MP::sub(t, x0);
        return MP::add(t, F);
    } else {
        double t = xp1 - xm2;
        t = MP::fma_(t, xm1, -x0);
        return t + F;
    }
}

// dx_i = (((-x_{i-2}) * x_{i-1} + x_{i-1} * x_{i+1}) - x_i) + F   (Euler plugin;
// reference tests/solvers/mock_fun3d/fun3d:54)
template <class MP>
FDS_HD double l96_rhs_fun3d(double xm2, double xm1, double x0, double xp1, double F) {
    if (MP::kExact) {
        double a = MP::mul(-xm2, xm1);
        double b = MP::mul(xm1, xp1);
        double t = MP::add(a, b);
        t = MP::sub(t, x0);
        return MP::add(t, F);
    } else {
        double b = xm1 * xp1;
        double t = MP::fma_(-xm2, xm1, b);
        return (t - x0) + F;
    }
}

struct Rk4Coef {
    double h2;  // 0.5*dt
    double h;   // dt
    double h6;  // dt/6.0
};

// --------------------------------------------------------------------------- //
// Software-pipelined RK4 along one column as a SKEWED wavefront: push x[i+4], get
// x'[i-3].  The four stages of one push work on different sites,
//     stage 1 at i+3 : k1 from x       -> y2[i+3], acc[i+3]  = k1
//     stage 2 at i+1 : k2 from y2      -> y3[i+1], acc[i+1] += 2 k2
//     stage 3 at i-1 : k3 from y3      -> y4[i-1], acc[i-1] += 2 k3
//     stage 4 at i-3 : k4 from y4      -> x'[i-3] = x + h6 (acc + k4)
// and each reads only values produced by EARLIER pushes, so the four dependency
// chains are independent and the FP64 pipe sees 4-way instruction-level
// parallelism per thread.  Rings are indexed by (site mod 8); P = i mod 8 is a
// template parameter, so every ring index is a constant and the rings live in
// registers (only the live entries: 8 + 5 + 5 + 5 + 7 doubles).
// --------------------------------------------------------------------------- //
template <class MP>
struct Rk4Pipe {
    double X[8], Y2[8], Y3[8], Y4[8], A[8];
    static constexpr int kLag = 3;      // output site = push site - kLag

    FDS_HD void reset() {
#pragma unroll
        for (int j = 0; j < 8; ++j) { X[j] = 0.0; Y2[j] = 0.0; Y3[j] = 0.0; Y4[j] = 0.0; A[j] = 0.0; }
    }

    template <int P>
    FDS_HD double push(double xn, double F, const Rk4Coef& c) {
#define FDS_SLOT(o) ((P + (o) + 16) & 7)
        X[FDS_SLOT(4)] = xn;
        // stage 4 at site i-3
        const double k4 = l96_rhs<MP>(Y4[FDS_SLOT(-5)], Y4[FDS_SLOT(-4)], Y4[FDS_SLOT(-3)], Y4[FDS_SLOT(-2)], F);
        const double out = MP::muladd(c.h6, MP::add(A[FDS_SLOT(-3)], k4), X[FDS_SLOT(-3)]);
        // stage 3 at site i-1
        const double k3 = l96_rhs<MP>(Y3[FDS_SLOT(-3)], Y3[FDS_SLOT(-2)], Y3[FDS_SLOT(-1)], Y3[FDS_SLOT(0)], F);
        A[FDS_SLOT(-1)] = MP::twice_plus(k3, A[FDS_SLOT(-1)]);
        Y4[FDS_SLOT(-1)] = MP::muladd(c.h, k3, X[FDS_SLOT(-1)]);
        // stage 2 at site i+1
        const double k2 = l96_rhs<MP>(Y2[FDS_SLOT(-1)], Y2[FDS_SLOT(0)], Y2[FDS_SLOT(1)], Y2[FDS_SLOT(2)], F);
        A[FDS_SLOT(1)] = MP::twice_plus(k2, A[FDS_SLOT(1)]);
        Y3[FDS_SLOT(1)] = MP::muladd(c.h2, k2, X[FDS_SLOT(1)]);
        // stage 1 at site i+3
        const double k1 = l96_rhs<MP>(X[FDS_SLOT(1)], X[FDS_SLOT(2)], X[FDS_SLOT(3)], X[FDS_SLOT(4)], F);
        A[FDS_SLOT(3)] = k1;
        Y2[FDS_SLOT(3)] = MP::muladd(c.h2, k1, X[FDS_SLOT(3)]);
#undef FDS_SLOT
        return out;
    }
};

// Forward Euler (the reference's own Lorenz-96 solver, tests/solvers/mock_fun3d/fun3d:48-57)
// behind the same interface: push x[i+4], get x'[i-3], so that strip plan, halo widths, IO
// policy and objective tree of the streaming kernel are shared with RK4.  x'[i-3] needs
// x[i-5 .. i-2]: the ring holds x[i-3 .. i+4], the two older values ride in two registers.
template <class MP>
struct EulerPipe {
    double X[8], xm4, xm5;
    static constexpr int kLag = 3;

    FDS_HD void reset() {
#pragma unroll
        for (int j = 0; j < 8; ++j) X[j] = 0.0;
        xm4 = 0.0; xm5 = 0.0;
    }

    template <int P>
    FDS_HD double push(double xn, double F, const Rk4Coef& c) {
#define FDS_SLOT(o) ((P + (o) + 16) & 7)
        xm5 = xm4;
        xm4 = X[FDS_SLOT(4)];              // the slot about to be overwritten holds x[i-4]
        X[FDS_SLOT(4)] = xn;
        const double k = l96_rhs_fun3d<MP>(xm5, xm4, X[FDS_SLOT(-3)], X[FDS_SLOT(-2)], F);
        return MP::muladd(c.h, k, X[FDS_SLOT(-3)]);
#undef FDS_SLOT
    }
};

// One RK4 step at a single site from its 13-point dependency cone
// x[0..12] = x_{i-8} .. x_{i+4}  (straightforward form; used by the simple kernel).
template <class MP>
FDS_HD double l96_rk4_point(const double* x, double F, const Rk4Coef& c) {
    double k1[10], y2[10];           // sites i-6 .. i+3   (x index j+2)
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        k1[j] = l96_rhs<MP>(x[j], x[j + 1], x[j + 2], x[j + 3], F);
        y2[j] = MP::muladd(c.h2, k1[j], x[j + 2]);
    }
    double k2[7], y3[7];             // sites i-4 .. i+2   (x index j+4)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        k2[j] = l96_rhs<MP>(y2[j], y2[j + 1], y2[j + 2], y2[j + 3], F);
        y3[j] = MP::muladd(c.h2, k2[j], x[j + 4]);
    }
    double k3[4], y4[4];             // sites i-2 .. i+1   (x index j+6)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        k3[j] = l96_rhs<MP>(y3[j], y3[j + 1], y3[j + 2], y3[j + 3], F);
        y4[j] = MP::muladd(c.h, k3[j], x[j + 6]);
    }
    const double k4 = l96_rhs<MP>(y4[0], y4[1], y4[2], y4[3], F);
    double s = MP::twice_plus(k2[4], k1[6]);
    s = MP::twice_plus(k3[2], s);
    s = MP::add(s, k4);
    return MP::muladd(c.h6, s, x[8]);
}

// forward Euler at a single site, x[0..3] = x_{i-2} .. x_{i+1}
template <class MP>
FDS_HD double l96_euler_point(const double* x, double F, double dt) {
    const double k = l96_rhs_fun3d<MP>(x[0], x[1], x[2], x[3], F);
    return MP::muladd(dt, k, x[2]);
}

// --------------------------------------------------------------------------- //
// Specified reduction tree (oracle/plugins.py: tree_sum): balanced binary tree
// over adjacent sites.  Tree8 folds 8 consecutive values; LeafStack folds 8
// consecutive Tree8 results into a 64-site leaf; TreeAcc is the generic
// binary-counter form.  Left operand is always the lower-index partial.
// --------------------------------------------------------------------------- //
struct Tree8 {
    double t1, t2, t4;
    template <int P>
    FDS_HD double push(double v) {   // returns the 8-sum when P == 7
        if (P == 0) { t1 = v; }
        else if (P == 1) { t2 = t1 + v; }
        else if (P == 2) { t1 = v; }
        else if (P == 3) { t4 = t2 + (t1 + v); }
        else if (P == 4) { t1 = v; }
        else if (P == 5) { t2 = t1 + v; }
        else if (P == 6) { t1 = v; }
        else { return t4 + (t2 + (t1 + v)); }
        return 0.0;
    }
};

struct LeafStack {
    double s0, s1, s2;
    // b = index of the 8-block inside its 64-site leaf; true when the leaf is complete.
    // Branch-free binary counter: the three sums are always formed and selected by the bits of b
    // (a sum whose operands are not both due is simply not used), so the hot loop has no
    // data-independent-but-unpredictable control flow and no register shuffling at merge points.
    FDS_HD bool push(double p, int b, double& leaf) {
        const bool b0 = (b & 1) != 0, b1 = (b & 2) != 0, b2 = (b & 4) != 0;
        const double t0 = s0 + p;
        const double p1 = b0 ? t0 : p;          // value carried into level 1 (meaningful when b0)
        s0 = b0 ? s0 : p;
        const double t1 = s1 + p1;
        const double p2 = b1 ? t1 : p1;         // carried into level 2 (meaningful when b0 && b1)
        s1 = (b0 && !b1) ? p1 : s1;
        const double t2 = s2 + p2;
        s2 = (b0 && b1 && !b2) ? p2 : s2;
        leaf = t2;
        return b0 && b1 && b2;
    }
    // The same counter without the leaf output: the eighth push (b == 7) leaves s0, s1, s2 untouched, so the
    // leaf sum can be formed LATER from the unchanged state by leaf(p) -- the hot loop then carries no branch
    // and no live leaf value between the push and the (rare) node update.
    FDS_HD void push_state(double p, int b) {
        const bool b0 = (b & 1) != 0, b1 = (b & 2) != 0, b2 = (b & 4) != 0;
        const double t0 = s0 + p;
        const double p1 = b0 ? t0 : p;
        s0 = b0 ? s0 : p;
        const double t1 = s1 + p1;
        const double p2 = b1 ? t1 : p1;
        s1 = (b0 && !b1) ? p1 : s1;
        s2 = (b0 && b1 && !b2) ? p2 : s2;
    }
    // The same with the lowest bit of b known at compile time (the unchecked loop of the two-step march handles two
    // blocks per iteration): an even block only parks its sum (no add, no select), an odd block pairs it with the
    // parked one and runs the two upper levels of the counter -- 2 adds and 3 selects instead of 2 and 5.
    template <int PAR>
    FDS_HD void push_state_p(double p, int b) {
        if (PAR == 0) { s0 = p; return; }
        const bool b1 = (b & 2) != 0, b2 = (b & 4) != 0;
        const double p1 = s0 + p;
        const double t1 = s1 + p1;
        const double p2 = b1 ? t1 : p1;
        s1 = b1 ? s1 : p1;
        s2 = (b1 && !b2) ? p2 : s2;
    }
    // leaf sum after the eighth push_state(p, 7): s2 + (s1 + (s0 + p)), the operation order of push()
    FDS_HD double leaf(double p) const { return s2 + (s1 + (s0 + p)); }
};

template <int MAXL>
struct TreeAcc {
    double st[MAXL + 1];
    unsigned cnt;
    FDS_HD void reset() { cnt = 0; }
    FDS_HD void push(double v) {
        int l = 0;
#pragma unroll 1
        while ((cnt >> l) & 1u) { v = st[l] + v; ++l; }
        st[l] = v;
        ++cnt;
    }
    // valid after exactly 2^L pushes
    FDS_HD double result(int L) const { return st[L]; }
};

constexpr int kLeafSites = 64;   // sites per leaf of the objective partials
constexpr int kLeafShift = 6;
constexpr int kMaxNodeLevels = 7;    // a node of the emitted partials spans at most 64 * 2^7 = 8192 sites

// Continuation of the tree from 64-site leaves up to nodes of 2^(6 + levels) sites, for the two
// objective sums at once: a binary counter whose pending partials live in storage provided by
// the IO policy (shared memory in the CUDA kernel -- it is touched once per 64 sites and
// indexed dynamically, so registers and local memory are both the wrong place).
//   double ns_get(int level, int which)      void ns_put(int level, int which, double v)
struct NodeCounter {
    unsigned cnt;
    FDS_HD void reset() { cnt = 0; }
    // push one leaf; true when a node of 2^levels leaves is complete (its sums in a, b)
    template <class IO>
    FDS_HD bool push(IO& io, double& a, double& b, int levels) {
        int l = 0;
#pragma unroll 1
        while ((cnt >> l) & 1u) { a = io.ns_get(l, 0) + a; b = io.ns_get(l, 1) + b; ++l; }
        ++cnt;
        if (l == levels) { cnt = 0; return true; }
        io.ns_put(l, 0, a); io.ns_put(l, 1, b);
        return false;
    }
};

// --------------------------------------------------------------------------- //
// Strip decomposition of the streaming step kernel.  FIXED per context: the widest
// output range a step can have, [-8(H-1), n + 4(H-1)) for a halo valid for H steps,
// is cut into strips of `ls` sites (a multiple of 64) anchored at a multiple of 64,
// so that every 8-block and every 64-site leaf is aligned to global site 0.  Every
// step computes the whole span (cells outside the still-valid halo receive values
// that are never used), so the kernel needs no range checks, and the arrays are
// allocated for sites [alo, ahi) = [a0 - 12, a0 + nstrips*ls + 12): exactly the
// input chunks the strips read (both bounds == 4 mod 8); the two-steps-per-pass kernel
// (l96_march2.cuh) starts one block earlier and drains one block later (20 instead of 12), a pass of K steps
// (l96_marchk.cuh) K - 1 blocks: the plan pads for kMaxFuse steps per pass.
// --------------------------------------------------------------------------- //
constexpr int kMaxFuse = 4;                          // most time steps one pass of the streaming kernel advances
constexpr int kPlanPad = 12 + 8 * (kMaxFuse - 1) + 8;   // sites the arrays extend beyond the strips on either side (+ 8: the
                                                       // two-block ring of the fused kernels fetches one block ahead of a strip's first)

struct StripPlan {
    long long a0;       // first strip starts here (multiple of the node size)
    long long ls;       // strip length (multiple of the node size)
    long long nstrips;  // strips needed to cover the span
    int nblk;           // 8-blocks per strip
    long long alo, ahi; // allocated site range
    int node_shift;     // objective partials are emitted per node of 2^node_shift sites (>= 6)
    long long nnode;    // nodes covering [0, n)
};

FDS_HD long long floor_div(long long a, long long b) {
    long long q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

inline StripPlan make_strip_plan(long long n, long long H, long long max_strips) {
    StripPlan p;
    if (max_strips < 1) max_strips = 1;
    const long long lo = -8 * (H - 1), hi = n + 4 * (H - 1);
    // the coarsest node (fewest partials to reduce afterwards) that lengthens a strip by <= 0.5 %
    // (the march is serial along a strip, so its length IS the kernel time: at N = 2^24, M = 34 nodes of 1024
    // sites round 14.7 nodes up to 15 = +1.7 %, nodes of 256 sites give the same length as 64-site leaves)
    long long ls6 = 0;
    for (int e = kLeafShift + kMaxNodeLevels; e >= kLeafShift; --e) {
        const long long g = 1LL << e;
        const long long a0 = floor_div(lo, g) * g;
        const long long span = hi - a0;
        long long ls = ((span + max_strips - 1) / max_strips + g - 1) / g * g;
        if (ls < g) ls = g;
        if (e > kLeafShift) {
            if (ls6 == 0) {   // reference length at the finest granularity
                const long long a6 = floor_div(lo, kLeafSites) * kLeafSites;
                const long long s6 = hi - a6;
                ls6 = ((s6 + max_strips - 1) / max_strips + kLeafSites - 1) / kLeafSites * kLeafSites;
                if (ls6 < kLeafSites) ls6 = kLeafSites;
            }
            if (ls * 200 > ls6 * 201) continue;
        }
        p.a0 = a0;
        p.ls = ls;
        p.nstrips = (span + ls - 1) / ls;
        p.node_shift = e;
        break;
    }
    p.nblk = (int)(p.ls / 8);
    p.alo = p.a0 - kPlanPad;
    p.ahi = p.a0 + p.nstrips * p.ls + kPlanPad;
    p.nnode = (n + (1LL << p.node_shift) - 1) >> p.node_shift;
    return p;
}

}  // namespace fds
