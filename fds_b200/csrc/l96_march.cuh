// The per-thread body of the streaming RK4 step: one thread marches along one
// strip of sites for one ensemble column.  Templated on an IO policy so that the
// CUDA kernel (shared-memory ring fed by TMA bulk copies) and the host emulation
// harness (plain arrays) execute literally the same code.
//
// IO policy:
//   void   acquire(int b)                 block b's 8 input sites are readable
//   double load(int b, int p)             x[a + 8b + p + 4]
//   void   release(int b)                 done reading block b
//   void   store(long long i, double v)   x'[i]
//   void   emit_leaf(long long leaf, double s1, double s2)   64-site partial sums
#pragma once
#include "l96_core.cuh"

namespace fds {

struct MarchGeom {
    long long a;      // first output site of this strip (multiple of 64)
    long long lo, hi; // outputs are stored for sites in [lo, hi)
    long long n;      // objective sums cover sites in [0, n)
    int nblk;         // 8-blocks in the strip
    int stats;        // accumulate the objective partials?
};

template <class MP, int P, class IO>
FDS_HD void march_site(IO& io, Rk4Pipe<MP>& pipe, Tree8& t1, Tree8& t2, double& b1, double& b2,
                       const MarchGeom& g, int b, long long i0, double F, const Rk4Coef& c,
                       bool live) {
    const double xn = io.load(b, P);
    const double xo = pipe.template push<P>(xn, F, c);
    const long long i = i0 + P;
    if (live && i >= g.lo && i < g.hi) io.store(i, xo);
    if (g.stats) {
        const double v = (i >= 0 && i < g.n) ? xo : 0.0;
        const double r1 = t1.template push<P>(v);
        const double r2 = t2.template push<P>(MP::mul(v, v));   // never contracted into the tree's add
        if (P == 7) { b1 = r1; b2 = r2; }
    }
}

template <class MP, class IO>
FDS_HD void l96_rk4_march(IO& io, const MarchGeom& g, double F, const Rk4Coef& c) {
    Rk4Pipe<MP> pipe;
    pipe.reset();
    Tree8 t1, t2;
    LeafStack l1, l2;
    t1.t1 = t1.t2 = t1.t4 = 0.0; t2 = t1;
    l1.s0 = l1.s1 = l1.s2 = 0.0; l2 = l1;
    // two warm-up blocks fill the stage pipeline (their outputs belong to the
    // previous strip and are discarded)
#pragma unroll 1
    for (int b = -2; b < g.nblk; ++b) {
        const long long i0 = g.a + 8LL * b;
        const bool live = b >= 0;
        double b1 = 0.0, b2 = 0.0;
        io.acquire(b);
        march_site<MP, 0>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 1>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 2>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 3>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 4>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 5>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 6>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        march_site<MP, 7>(io, pipe, t1, t2, b1, b2, g, b, i0, F, c, live);
        io.release(b);
        if (g.stats && live) {
            double s1, s2;
            const bool d1 = l1.push(b1, b & 7, s1);
            const bool d2 = l2.push(b2, b & 7, s2);
            if (d1 && d2) io.emit_leaf(i0 >> kLeafShift, s1, s2);
        }
    }
}

}  // namespace fds
