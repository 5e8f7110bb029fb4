// Host emulation of the streaming RK4 step kernel's per-thread control flow.
// Built with g++ -ffp-contract=off into tests/_emul.so and compared bit for bit
// with the numpy oracle on the CPU (tests/test_emulation.py).  It executes the
// same headers (l96_core.cuh, l96_march.cuh) as the CUDA kernel; only the IO
// policy (plain arrays instead of a TMA-fed shared-memory ring) differs.
#include <vector>
#include "l96_march.cuh"
#include "l96_march2.cuh"
#include "l96_marchk.cuh"

using namespace fds;

namespace {
struct HostIO {
    const double* X;   // base of site 0 of the padded input, [site][M]
    double* Y;
    double* jpart;     // [nnode][M][2]
    long long nleaf;   // = nnode
    long long a;
    int M, k;
    long long cs = 0;     // first input site of the current block
    long long j0 = 0;     // first output site of the current block
    void begin_block(int b) {
        cs = a + 8LL * b + 4;
        j0 = a + 8LL * b - 3;
    }
    void end_block(int) {}
    double load(int p) const { return X[(cs + p) * M + k]; }
    void store(int p, double v) { Y[(j0 + p) * M + k] = v; }
    void emit_node(long long node, double s1, double s2) {
        if (node >= 0 && node < nleaf) {
            jpart[(node * M + k) * 2 + 0] = s1;
            jpart[(node * M + k) * 2 + 1] = s2;
        }
    }
    bool all(bool v) const { return v; }
    double ns[kMaxNodeLevels + 1][2];
    double ns_get(int l, int q) const { return ns[l][q]; }
    void ns_put(int l, int q, double v) { ns[l][q] = v; }
};
// two steps per pass: output cursor at site a + 8b - 11, two sets of objective partials
struct HostIO2 {
    const double* X;
    double* Y;
    double* jpart[2];  // [nnode][M][2] of step 1 and step 2
    long long nleaf;
    long long a;
    int M, k;
    long long cs = 0, j0 = 0;
    template <int H = -1>
    void begin_block(int b) {
        cs = a + 8LL * b + 4;
        j0 = a + 8LL * b - 11;
    }
    template <int H = -1>
    void end_block(int) {}
    double load(int p) const { return X[(cs + p) * M + k]; }
    void store(int p, double v) { Y[(j0 + p) * M + k] = v; }
    void emit_node2(int step, long long node, double s1, double s2) {
        if (node >= 0 && node < nleaf) {
            jpart[step][(node * M + k) * 2 + 0] = s1;
            jpart[step][(node * M + k) * 2 + 1] = s2;
        }
    }
    bool all(bool v) const { return v; }
    template <int H = -1>
    void mid_block() {}
    int warp_max(int v) const { return v; }
    int warp_min(int v) const { return v; }
    double ns[2][kMaxNodeLevels + 1][2];
    double ns_get2(int st, int l, int q) const { return ns[st][l][q]; }
    void ns_put2(int st, int l, int q, double v) { ns[st][l][q] = v; }
    double pl[2 * kMaxFuse][4];
    void pl_put(int sum, int slot, double v) { pl[sum][slot] = v; }
    double pl_get(int sum, int slot) const { return pl[sum][slot]; }
    int plc = 0;
    void pl_cur_set(int slot) { plc = slot; }
    void pl_cur_next() { ++plc; }
    template <int SUM>
    void pl_cur_put(double v) { pl[SUM][plc] = v; }
};
// K steps per pass (l96_marchk.cuh): output cursor at site a + 8b - 3 - 8(K-1), K sets of objective partials
template <int K>
struct HostIOK {
    const double* X;
    double* Y;
    double* const* jpart;   // [K] -> [nnode][M][2]
    long long nleaf;
    long long a;
    int M, k;
    long long cs = 0, j0 = 0;
    template <int H = -1>
    void begin_block(int b) {
        cs = a + 8LL * b + 4;
        j0 = a + 8LL * b - 3 - 8LL * (K - 1);
    }
    template <int H = -1>
    void end_block(int) {}
    double load(int p) const { return X[(cs + p) * M + k]; }
    void store(int p, double v) { Y[(j0 + p) * M + k] = v; }
    void emit_node2(int step, long long node, double s1, double s2) {
        if (node >= 0 && node < nleaf) {
            jpart[step][(node * M + k) * 2 + 0] = s1;
            jpart[step][(node * M + k) * 2 + 1] = s2;
        }
    }
    bool all(bool v) const { return v; }
    template <int H = -1>
    void mid_block() {}
    int warp_max(int v) const { return v; }
    int warp_min(int v) const { return v; }
    double ns[K][kMaxNodeLevels + 1][2];
    double ns_get2(int st, int l, int q) const { return ns[st][l][q]; }
    void ns_put2(int st, int l, int q, double v) { ns[st][l][q] = v; }
    double pl[2 * kMaxFuse][4];
    void pl_put(int sum, int slot, double v) { pl[sum][slot] = v; }
    double pl_get(int sum, int slot) const { return pl[sum][slot]; }
    int plc = 0;
    void pl_cur_set(int slot) { plc = slot; }
    void pl_cur_next() { ++plc; }
    template <int SUM>
    void pl_cur_put(double v) { pl[SUM][plc] = v; }
};
}  // namespace

extern "C" {

// the fixed strip plan of a context: out = {a0, ls, nstrips, nblk, alo, ahi, node_shift, nnode}
int emul_plan(long long n, long long H, long long max_strips, long long* out) {
    StripPlan p = make_strip_plan(n, H, max_strips);
    out[0] = p.a0; out[1] = p.ls; out[2] = p.nstrips; out[3] = p.nblk; out[4] = p.alo; out[5] = p.ahi;
    out[6] = p.node_shift; out[7] = p.nnode;
    return 0;
}

// One step over the whole span of the plan.  X, Y: pointers to site 0 of padded [site][M]
// arrays allocated for sites [alo, ahi) of the plan.
int emul_l96_rk4_step(const double* X, double* Y, int M, long long n, long long H,
                      const double* F /*[M]*/, double dt, int exact, long long max_strips, double* jpart) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIO io{X, Y, jpart, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jpart != nullptr, plan.node_shift};
            if (exact) l96_rk4_march<ExactMath>(io, g, F[k], c);
            else       l96_rk4_march<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// TWO steps in one pass (l96_march2.cuh); jpart1 / jpart2: objective partials of step 1 and step 2
int emul_l96_rk4_step2(const double* X, double* Y, int M, long long n, long long H,
                       const double* F /*[M]*/, double dt, int exact, long long max_strips, double* jpart1,
                       double* jpart2) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIO2 io{X, Y, {jpart1, jpart2}, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jpart1 != nullptr, plan.node_shift};
            if (exact) l96_rk4_march2<ExactMath>(io, g, F[k], c);
            else       l96_rk4_march2<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// two forward-Euler steps in one pass
int emul_l96_euler_step2(const double* X, double* Y, int M, long long n, long long H,
                         const double* F /*[M]*/, double dt, int exact, long long max_strips, double* jpart1,
                         double* jpart2) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIO2 io{X, Y, {jpart1, jpart2}, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jpart1 != nullptr, plan.node_shift};
            if (exact) l96_euler_march2<ExactMath>(io, g, F[k], c);
            else       l96_euler_march2<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// FOUR forward-Euler steps in one pass (l96_marchk.cuh); jparts: 4 pointers
int emul_l96_euler_step4(const double* X, double* Y, int M, long long n, long long H,
                         const double* F /*[M]*/, double dt, int exact, long long max_strips, double* const* jparts) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIOK<4> io{X, Y, jparts, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jparts != nullptr, plan.node_shift};
            if (exact) l96_euler_march4<ExactMath>(io, g, F[k], c);
            else       l96_euler_march4<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// the generic K-step march with K = 2 and the RK4 pipe (must equal emul_l96_rk4_step2)
int emul_l96_rk4_stepk2(const double* X, double* Y, int M, long long n, long long H,
                        const double* F /*[M]*/, double dt, int exact, long long max_strips, double* const* jparts) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIOK<2> io{X, Y, jparts, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jparts != nullptr, plan.node_shift};
            if (exact) l96_rk4_marchk2<ExactMath>(io, g, F[k], c);
            else       l96_rk4_marchk2<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// the same with the forward-Euler pipe
int emul_l96_euler_step(const double* X, double* Y, int M, long long n, long long H,
                        const double* F /*[M]*/, double dt, int exact, long long max_strips, double* jpart) {
    StripPlan plan = make_strip_plan(n, H, max_strips);
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long s = 0; s < plan.nstrips; ++s) {
        for (int k = 0; k < M; ++k) {
            HostIO io{X, Y, jpart, plan.nnode, plan.a0 + s * plan.ls, M, k};
            MarchGeom g{io.a, n, plan.nblk, jpart != nullptr, plan.node_shift};
            if (exact) l96_euler_march<ExactMath>(io, g, F[k], c);
            else       l96_euler_march<FastMath>(io, g, F[k], c);
        }
    }
    return 0;
}

// simple-kernel math: one site from its 13-point cone
int emul_l96_rk4_point(const double* X, double* Y, int M, long long lo, long long hi,
                       const double* F, double dt, int exact) {
    Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    for (long long i = lo; i < hi; ++i)
        for (int k = 0; k < M; ++k) {
            double x[13];
            for (int j = 0; j < 13; ++j) x[j] = X[(i - 8 + j) * M + k];
            Y[i * M + k] = exact ? l96_rk4_point<ExactMath>(x, F[k], c) : l96_rk4_point<FastMath>(x, F[k], c);
        }
    return 0;
}

int emul_l96_euler_point(const double* X, double* Y, int M, long long lo, long long hi,
                         const double* F, double dt, int exact) {
    for (long long i = lo; i < hi; ++i)
        for (int k = 0; k < M; ++k) {
            double x[4];
            for (int j = 0; j < 4; ++j) x[j] = X[(i - 2 + j) * M + k];
            Y[i * M + k] = exact ? l96_euler_point<ExactMath>(x, F[k], dt) : l96_euler_point<FastMath>(x, F[k], dt);
        }
    return 0;
}

// canonical tree over `count` leaf partials (padded with zeros to a power of two)
double emul_tree_reduce(const double* v, long long count, long long stride) {
    long long p = 1; int L = 0;
    while (p < count) { p *= 2; ++L; }
    TreeAcc<40> acc; acc.reset();
    for (long long j = 0; j < p; ++j) acc.push(j < count ? v[j * stride] : 0.0);
    return acc.result(L);
}

}  // extern "C"
