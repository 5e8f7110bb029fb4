// The per-thread body of the streaming RK4 step: one thread marches along one
// strip of sites for one ensemble column.  Templated on an IO policy so that the
// CUDA kernel (shared-memory ring fed by TMA bulk copies) and the host emulation
// harness (plain arrays) execute literally the same code.
//
// Block b of a strip that starts at site a consumes the 8 input sites
// a + 8b + 4 .. a + 8b + 11 and produces the 8 output sites a + 8b - 3 .. a + 8b + 4
// (the pipeline lags 3 sites behind the push index, l96_core.cuh).
//
// IO policy:
//   void   begin_block(int b)             block b's 8 input sites are readable; position the
//                                         output cursor at site a + 8b - 3
//   double load(int p)                    x[a + 8b + 4 + p]
//   void   store(int p, double v)         x'[a + 8b - 3 + p]
//   void   end_block(int b)               done with block b
//   void   emit_node(long long node, double s1, double s2)   partial sums of one node (2^node_shift sites)
//   double ns_get(int level, int which) / void ns_put(int level, int which, double v)
//                                         per-thread storage of the node counter (l96_core.cuh)
//   bool   all(bool)                      warp-wide AND (the CUDA kernel keeps a warp on ONE
//                                         of the two code paths; true for a single thread)
#pragma once
#include "l96_core.cuh"

namespace fds {

struct MarchGeom {
    long long a;      // first output site of this strip (multiple of 64)
    long long n;      // objective sums cover sites in [0, n)
    int nblk;         // 8-blocks in the strip; the strip stores sites [a, a + 8*nblk)
    int stats;        // accumulate the objective partials?
    int node_shift;   // partials are emitted per node of 2^node_shift sites (a, 8*nblk multiples of it)
};

// One site of one block.  FAST: the whole block lies inside the store range and the
// objective range, so no per-site checks are needed.
template <class MP, class PIPE, int P, bool FAST, bool STATS, class IO>
FDS_HD void march_site(IO& io, PIPE& pipe, Tree8& t1, Tree8& t2, double& b1, double& b2,
                       long long j0, long long slo, long long shi, long long jlo, long long jhi,
                       double F, const Rk4Coef& c) {
    constexpr bool stats = STATS;
    constexpr int Q = (P + 8 - PIPE::kLag) & 7;   // output site mod 8
    const double xo = pipe.template push<P>(io.load(P), F, c);
    if (FAST) {
        io.store(P, xo);
        if (stats) {
            const double r1 = t1.template push<Q>(xo);
            const double r2 = t2.template push<Q>(MP::mul(xo, xo));   // never contracted into the tree's add
            if (Q == 7) { b1 = r1; b2 = r2; }
        }
    } else {
        const long long j = j0 + P;
        if (j >= slo && j < shi) io.store(P, xo);
        if (stats) {
            const double v = (j >= jlo && j < jhi) ? xo : 0.0;
            const double r1 = t1.template push<Q>(v);
            const double r2 = t2.template push<Q>(MP::mul(v, v));
            if (Q == 7) { b1 = r1; b2 = r2; }
        }
    }
}

template <class MP, class PIPE, bool FAST, bool STATS, class IO>
FDS_HD void march_block(IO& io, PIPE& pipe, Tree8& t1, Tree8& t2, LeafStack& l1, LeafStack& l2,
                        NodeCounter& ns, const MarchGeom& g, int b, long long slo, long long shi, long long jlo, long long jhi,
                        double F, const Rk4Coef& c) {
    const long long j0 = g.a + 8LL * b - PIPE::kLag;
    constexpr bool stats = STATS;
    double b1 = 0.0, b2 = 0.0;
    march_site<MP, PIPE, 0, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 1, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 2, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    // sites a+8(b-1) .. a+8b-1 are complete here: 8-block (b-1) of the strip
    if (stats && b >= 1) {
        double s1, s2;
        const bool d1 = l1.push(b1, (b - 1) & 7, s1);
        const bool d2 = l2.push(b2, (b - 1) & 7, s2);
        if (d1 && d2) {
            // leaves outside [0, n) hold unused cells: they enter the tree as exact zeros
            const long long leaf = (g.a + 8LL * (b - 1)) >> kLeafShift;
            if (leaf < 0 || leaf * kLeafSites >= g.n) { s1 = 0.0; s2 = 0.0; }
            if (ns.push(io, s1, s2, g.node_shift - kLeafShift)) io.emit_node(leaf >> (g.node_shift - kLeafShift), s1, s2);
        }
    }
    march_site<MP, PIPE, 3, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 4, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 5, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 6, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
    march_site<MP, PIPE, 7, FAST, STATS>(io, pipe, t1, t2, b1, b2, j0, slo, shi, jlo, jhi, F, c);
}

// number of blocks a strip of nblk 8-blocks runs through: two warm-up blocks fill the
// stage pipeline, one trailing block drains the 3-site lag
FDS_HD int march_iterations(int nblk) { return nblk + 3; }

// STATS (compile time) = accumulate the objective partials; g.stats is ignored here
template <class MP, bool STATS, class PIPE, class IO>
FDS_HD void l96_march_t(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    PIPE pipe;
    pipe.reset();
    Tree8 t1, t2;
    LeafStack l1, l2;
    t1.t1 = t1.t2 = t1.t4 = 0.0; t2 = t1;
    l1.s0 = l1.s1 = l1.s2 = 0.0; l2 = l1;
    NodeCounter ns;
    ns.reset();
    // this strip stores sites [slo, shi) and counts sites [jlo, jhi) in the objective
    const long long slo = g.a, shi = g.a + 8LL * g.nblk;
    const long long jlo = slo > 0 ? slo : 0;
    const long long jhi = g.n < shi ? g.n : shi;
    // Only the 64-site leaf that contains site n (if n is not a multiple of 64) needs its
    // objective terms masked: leaves wholly outside [0, n) are zeroed as a whole (march_block).
    const long long nfl = g.n & ~(long long)(kLeafSites - 1);
    const long long ncl = (g.n + kLeafSites - 1) & ~(long long)(kLeafSites - 1);
#pragma unroll 1
    for (int b = -2; b <= g.nblk; ++b) {
        const long long j0 = g.a + 8LL * b - PIPE::kLag;
        io.begin_block(b);
        // unchecked path: all 8 outputs belong to this strip and none lies in the partial leaf
        const bool fast = j0 >= slo && j0 + 8 <= shi && (j0 + 8 <= nfl || j0 >= ncl);
        if (io.all(fast))
            march_block<MP, PIPE, true, STATS>(io, pipe, t1, t2, l1, l2, ns, g, b, slo, shi, jlo, jhi, F, c);
        else
            march_block<MP, PIPE, false, STATS>(io, pipe, t1, t2, l1, l2, ns, g, b, slo, shi, jlo, jhi, F, c);
        io.end_block(b);
    }
}

template <class MP, bool STATS, class IO>
FDS_HD void l96_rk4_march_t(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    l96_march_t<MP, STATS, Rk4Pipe<MP>>(io, g, F, c);
}

template <class MP, class IO>
FDS_HD void l96_rk4_march(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_march_t<MP, true, Rk4Pipe<MP>>(io, g, F, c);
    else l96_march_t<MP, false, Rk4Pipe<MP>>(io, g, F, c);
}

template <class MP, class IO>
FDS_HD void l96_euler_march(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    if (g.stats) l96_march_t<MP, true, EulerPipe<MP>>(io, g, F, c);
    else l96_march_t<MP, false, EulerPipe<MP>>(io, g, F, c);
}

}  // namespace fds
