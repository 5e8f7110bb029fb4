// Segment-boundary kernels: time-dilation vector, extended Gram, CholeskyQR2 small
// algebra, tangent update / next-ensemble construction.
//
// K3-K6 and K9 of SURVEY.md 2.3.  Reference: fds/timedilation.py:14-82 (dxdt,
// contribution, project), fds/lsstan.py:20-34 (checkpoint: QR, b, v update),
// fds/segment.py:43-51 (perturbed initial conditions), :74-79 (FD tangents).
//
// Tangents are rows of T[site][m+1] (j < m homogeneous, j = m inhomogeneous) or,
// right after a segment, implicit in the ensemble: t_j = (E[.][j+1] - E[.][0]) / eps.
// Everything at a boundary is linear in X_i = [t_0 .. t_m, d_i] (MX = m+2 values per
// site), so it reduces to (1) the Gram X^T X, one pass, (2) tiny dense algebra on that
// Gram in one CTA, (3) out_i = A X_i with a constant (m+1) x (m+2) matrix, one pass.
#include <algorithm>
#include "kernels.h"
#include "common.cuh"

namespace fds {

// ------------------------------ elementwise bits ----------------------------- //
__global__ void extract_col0_kernel(const double* __restrict__ E, int M, double* __restrict__ P, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P[i] = E[i * M];
}
cudaError_t launch_extract_col0(const double* E, int M, double* P, long long n, cudaStream_t st, LaunchStats* ls) {
    extract_col0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, M, P, n);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// d = ((c0*u0 + c1*u1) + c2*u2) + c3*u3, each product and sum rounded
// (fds/timedilation.py:20: Python sum() of c[i]*u[i], left to right)
__global__ void dxdt_kernel(const double* __restrict__ P0, const double* __restrict__ P1,
                            const double* __restrict__ P2, const double* __restrict__ P3,
                            const double* __restrict__ c, double* __restrict__ d, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = __dmul_rn(c[0], P0[i]);
    acc = __dadd_rn(acc, __dmul_rn(c[1], P1[i]));
    acc = __dadd_rn(acc, __dmul_rn(c[2], P2[i]));
    acc = __dadd_rn(acc, __dmul_rn(c[3], P3[i]));
    d[i] = acc;
}
cudaError_t launch_dxdt(const double* P0, const double* P1, const double* P2, const double* P3,
                        const double* c, double* d, long long n, cudaStream_t st, LaunchStats* ls) {
    dxdt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P0, P1, P2, P3, c, d, n);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// ---- time dilation, fused: u_1 .. u_order and d in one pass (see kernels.h) ---------------------------------- //
// A CTA owns a WINDOW of kDilThreads * kDilRun consecutive sites, thread t the run [t kDilRun, (t + 1) kDilRun) of
// it, in registers, for every stage of every step: the state u, the k-accumulator and d never touch memory.  A stage
// hands its input (u, y2, y3, y4) to the neighbouring threads through shared memory -- one store per site, three
// halo loads per thread and stage -- so the kernel is bound by neither shared-memory bandwidth (the stage-array form:
// 9 accesses per site and stage, 0.42 ms at N = 2^24) nor the FP64 pipe (the single-site form l96_rk4_point evaluates
// 22 right-hand sides per site and step: 0.46 ms).  The dependency cone eats 8 sites on the left and 4 on the right
// per step, so a window of W sites yields d on its inner W - 12 * order sites.  Same operations in the same order as
// l96_rk4_point / Rk4Pipe (one right-hand side per site, stage and step): bit-identical.
constexpr int kDilThreads = 256;
constexpr int kDilRun = 8;
constexpr int kDilWindow = kDilThreads * kDilRun;

//
// STATE = 1 is the same march used as the single-column solver (run-up, fds/fds.py:207): it writes the state after
// `order` steps for the sites [lo, hi) -- lo < 0 / hi > n reach into the halo cells of the array, which a multi-step
// advance consumes at 8 sites on the left and 4 on the right per step like the one-step kernel does -- instead of d.
// STATE = 2 also records the objective of every step (the solver call u1, J = run(u0, s, steps)): the windows then
// start at multiples of 64 sites, a thread's run is an aligned group of 8, eight neighbouring lanes a 64-site leaf of
// the canonical tree, and the leaf sums [sum x, sum x^2] of step k go to jp[k - 1][leaf][2] (sites >= sites count as
// zero, as in stats_leaf_kernel); the levels above the leaves are reduced by tree_level_kernel, all steps at once.
struct DilationGeom {
    int glw, tile;             // window positions [glw, glw + tile) are the ones this CTA writes
    long long lo, hi;          // sites written: [lo, hi)
    long long t00, rlo, rhi;   // first site of CTA 0's tile; sites that may be read: [rlo, rhi), zero outside
    long long sites, nleaf;    // STATE = 2: sites of the state proper, 64-site leaves
};

template <class MP, int EULER, int STATE>
__global__ void __launch_bounds__(kDilThreads) dilation_fused_kernel(const double* __restrict__ P0, double* __restrict__ d,
                                                                     DilationGeom g, DilationStencil st, double F0,
                                                                     Rk4Coef c, const double2* __restrict__ Fpp, long long slot,
                                                                     double* __restrict__ jp) {
    constexpr int L = kDilRun, W = kDilWindow;
    // window position j lives at sm[.][j + (j >> 3) + 3]: one unused slot per run of 8, so that the lanes of a warp
    // (runs 8 positions apart) hit different banks -- with the plain layout every exchange was an 8-way bank conflict
    // and the kernel spent its time in the shared-memory pipe (0.37 ms at N = 2^24)
    constexpr int SMW = W + W / L + 4;
    __shared__ __align__(16) double sm[2][SMW];
    const int p = st.order;
    const int tile = g.tile;
    const long long n = g.hi;
    const long long t0 = g.t00 + (long long)blockIdx.x * tile;          // first site this CTA writes d for
    const long long w0 = t0 - g.glw;                                    // site of window position 0
    const int tid = threadIdx.x, j0 = tid * L;
    if (tid < 2) { sm[0][tid] = 0.0; sm[1][tid] = 0.0; }               // positions -2, -1
    if (tid == 2) { sm[0][W + W / L + 3] = 0.0; sm[1][W + W / L + 3] = 0.0; }   // position W
    double u[L], dacc[L], F[L];
#pragma unroll
    for (int q = 0; q < L; ++q) {
        const long long site = w0 + j0 + q;
        u[q] = site >= g.rlo && site < g.rhi ? P0[site] : 0.0;
        dacc[q] = STATE ? 0.0 : __dmul_rn(st.c[0], u[q]);
        F[q] = F0;
        if (Fpp != nullptr) {
            long long pr = site / slot;                                  // ghosts left of stacked site 0 belong to problem 0
            F[q] = Fpp[pr < 0 ? 0 : pr].x;
        }
    }
    int buf = 0;
    // x[0 .. L+2] = stage input at window positions j0 - 2 .. j0 + L: the own run from registers, the halo from the
    // neighbours' stores
    auto exchange = [&](const double (&own)[L], double (&x)[L + 3]) {
        double* row = sm[buf] + (L + 1) * tid + 3;                      // position j0
#pragma unroll
        for (int q = 0; q < L; ++q) row[q] = own[q];
        __syncthreads();
        x[0] = row[-3]; x[1] = row[-2]; x[L + 2] = row[L + 1];           // positions j0 - 2, j0 - 1 (previous run), j0 + L (next run)
#pragma unroll
        for (int q = 0; q < L; ++q) x[q + 2] = own[q];
        buf ^= 1;                                                        // the next stage writes the other buffer: one barrier per stage
    };
    for (int k = 1; k <= p; ++k) {
        double x[L + 3];
        if (EULER) {
            exchange(u, x);
#pragma unroll
            for (int q = 0; q < L; ++q) u[q] = l96_euler_point<MP>(x + q, F[q], c.h);
        } else {
            double acc[L], y[L];
            exchange(u, x);
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const double k1 = l96_rhs<MP>(x[q], x[q + 1], x[q + 2], x[q + 3], F[q]);
                acc[q] = k1;
                y[q] = MP::muladd(c.h2, k1, u[q]);
            }
            exchange(y, x);
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const double k2 = l96_rhs<MP>(x[q], x[q + 1], x[q + 2], x[q + 3], F[q]);
                acc[q] = MP::twice_plus(k2, acc[q]);
                y[q] = MP::muladd(c.h2, k2, u[q]);
            }
            exchange(y, x);
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const double k3 = l96_rhs<MP>(x[q], x[q + 1], x[q + 2], x[q + 3], F[q]);
                acc[q] = MP::twice_plus(k3, acc[q]);
                y[q] = MP::muladd(c.h, k3, u[q]);
            }
            exchange(y, x);
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const double k4 = l96_rhs<MP>(x[q], x[q + 1], x[q + 2], x[q + 3], F[q]);
                u[q] = MP::muladd(c.h6, MP::add(acc[q], k4), u[q]);
            }
        }
        if (!STATE) {
#pragma unroll
            for (int q = 0; q < L; ++q) dacc[q] = __dadd_rn(dacc[q], __dmul_rn(st.c[k], u[q]));
        }
        if (STATE == 2) {
            // canonical tree over the 8 sites of the run, then over the 8 runs of the leaf (adjacent pairs first)
            double a[L], b[L];
#pragma unroll
            for (int q = 0; q < L; ++q) {
                const long long site = w0 + j0 + q;
                a[q] = site >= 0 && site < g.sites ? u[q] : 0.0;
                b[q] = __dmul_rn(a[q], a[q]);
            }
#pragma unroll
            for (int h = 1; h < L; h <<= 1) {
#pragma unroll
                for (int q = 0; q < L; q += 2 * h) { a[q] = __dadd_rn(a[q], a[q + h]); b[q] = __dadd_rn(b[q], b[q + h]); }
            }
#pragma unroll
            for (int h = 1; h < 8; h <<= 1) {
                a[0] = __dadd_rn(a[0], __shfl_xor_sync(0xffffffffu, a[0], h));
                b[0] = __dadd_rn(b[0], __shfl_xor_sync(0xffffffffu, b[0], h));
            }
            const long long site0 = w0 + j0;
            if ((tid & 7) == 0 && j0 >= g.glw && j0 < g.glw + tile && site0 >= 0 && site0 < g.sites)
                reinterpret_cast<double2*>(jp)[(size_t)(k - 1) * g.nleaf + (site0 >> 6)] = make_double2(a[0], b[0]);
        }
    }
#pragma unroll
    for (int q = 0; q < L; ++q) {
        const int j = j0 + q;
        const long long site = w0 + j;
        if (j >= g.glw && j < g.glw + tile && site >= g.lo && site < n) d[site] = STATE ? u[q] : dacc[q];
    }
}

cudaError_t launch_dilation_fused(const double* P0, double* d, long long n, const DilationStencil& st, double F0, double dt,
                                  int exact, int euler, const double2* Fpp, long long slot, cudaStream_t stream,
                                  LaunchStats* ls) {
    if (st.order < 1 || st.order > kMaxDilationOrder || n < 1) return cudaErrorInvalidValue;
    const int tile = kDilWindow - st.order * (euler ? 3 : 12);
    const unsigned grid = (unsigned)((n + tile - 1) / tile);
    const Rk4Coef c{0.5 * dt, dt, dt / 6.0};
    // the left halo always exists: the lowest site read is -order * GL
    const DilationGeom g{st.order * (euler ? 2 : 8), tile, 0, n, 0, -(long long)st.order * (euler ? 2 : 8),
                         n + (long long)st.order * (euler ? 1 : 4), 0, 0};
#define FDS_DIL(MP, E) dilation_fused_kernel<MP, E, 0><<<grid, kDilThreads, 0, stream>>>(P0, d, g, st, F0, c, Fpp, slot, nullptr)
    if (exact) { if (euler) FDS_DIL(ExactMath, 1); else FDS_DIL(ExactMath, 0); }
    else       { if (euler) FDS_DIL(FastMath, 1);  else FDS_DIL(FastMath, 0); }
#undef FDS_DIL
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// Y[lo, hi) = X advanced by `steps` (1 .. kMaxDilationOrder) time steps; X must be valid on
// [lo - 8 steps, hi + 4 steps) (2 / 1 per step for forward Euler).
// jp != nullptr: also the 64-site leaf sums of the objective of every step, jp[steps][nleaf][2], for the `sites`
// sites of the state proper (lo <= 0, hi >= sites).
// Fpp != nullptr: a batch -- [lo, hi) = all stacked sites, site i belongs to problem i / slot and takes its forcing;
// nothing is read below lo (the cells that would need it are ghost cells of problem 0 that the march gives up anyway).
cudaError_t launch_primal_multi(const double* X, double* Y, long long lo, long long hi, int steps, double F0, double dt,
                                int exact, int euler, double* jp, long long sites, long long nleaf,
                                const double2* Fpp, long long slot, cudaStream_t stream, LaunchStats* ls) {
    if (steps < 1 || steps > kMaxDilationOrder) return cudaErrorInvalidValue;
    if (hi <= lo) return cudaSuccess;
    if (jp != nullptr && (lo > 0 || hi < sites)) return cudaErrorInvalidValue;
    DilationStencil st{};
    st.order = steps;
    const int gl = euler ? 2 : 8, gr = euler ? 1 : 4;
    DilationGeom g{};
    g.lo = lo; g.hi = hi; g.sites = sites; g.nleaf = nleaf;
    g.rlo = Fpp ? lo : lo - (long long)steps * gl;
    g.rhi = Fpp ? hi : hi + (long long)steps * gr;
    if (jp != nullptr && Fpp != nullptr) return cudaErrorInvalidValue;
    if (jp != nullptr) {
        // leaf-aligned windows: margins of 64 sites (>= 8 steps * 8 and 8 steps * 4) and a whole number of leaves
        g.glw = 64;
        g.tile = kDilWindow - 128;
        g.t00 = -(((-lo) + 63) / 64) * 64;
    } else {
        g.glw = steps * gl;
        g.tile = kDilWindow - steps * (gl + gr);
        g.t00 = lo;
    }
    const unsigned grid = (unsigned)((hi - g.t00 + g.tile - 1) / g.tile);
    const Rk4Coef c{0.5 * dt, dt, dt / 6.0};
#define FDS_PRI(MP, E)                                                                                              \
    do { if (jp) dilation_fused_kernel<MP, E, 2><<<grid, kDilThreads, 0, stream>>>(X, Y, g, st, F0, c, nullptr, 1, jp); \
         else dilation_fused_kernel<MP, E, 1><<<grid, kDilThreads, 0, stream>>>(X, Y, g, st, F0, c, Fpp, Fpp ? slot : 1, nullptr); } while (0)
    if (exact) { if (euler) FDS_PRI(ExactMath, 1); else FDS_PRI(ExactMath, 0); }
    else       { if (euler) FDS_PRI(FastMath, 1);  else FDS_PRI(FastMath, 0); }
#undef FDS_PRI
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// T[i][j] = (E[i][j+1] - E[i][0]) / eps   (fds/segment.py:74,77; true IEEE division)
__global__ void diff_tangents_kernel(const double* __restrict__ E, int M, double eps, double* __restrict__ T,
                                     long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int MO = M - 1;
    const long long i = idx / MO;
    const int j = (int)(idx % MO);
    T[idx] = __ddiv_rn(__dsub_rn(E[i * M + j + 1], E[i * M]), eps);
}
cudaError_t launch_diff_tangents(const double* E, int M, double eps, double* T, long long n, cudaStream_t st,
                                 LaunchStats* ls) {
    const long long total = n * (M - 1);
    diff_tangents_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(E, M, eps, T, total);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// E[i][0] = u_i ; E[i][j+1] = u_i + T[i][j]*eps   (fds/segment.py:43,49; mul then add)
__global__ void make_ensemble_kernel(const double* __restrict__ u, const double* __restrict__ T, int M, double eps,
                                     double* __restrict__ E, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long i = idx / M;
    const int k = (int)(idx % M);
    const double ui = u[i];
    E[idx] = k == 0 ? ui : __dadd_rn(ui, __dmul_rn(T[i * (M - 1) + k - 1], eps));
}
cudaError_t launch_make_ensemble(const double* u, const double* T, int M, double eps, double* E, long long n,
                                 cudaStream_t st, LaunchStats* ls) {
    const long long total = n * M;
    make_ensemble_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(u, T, M, eps, E, total);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// tiled transposes between device [rows][cols] and host-order [cols][ld]
__global__ void transpose_kernel(const double* __restrict__ in, long long rows, int cols, double* __restrict__ out,
                                 long long ld_out) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long r = r0 + y;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[y][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int c = c0 + y;
        const long long r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(long long)c * ld_out + r] = tile[threadIdx.x][y];
    }
}
cudaError_t launch_transpose(const double* in, long long rows, int cols, double* out, long long ld_out,
                             cudaStream_t st, LaunchStats* ls) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(in, rows, cols, out, ld_out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}
__global__ void transpose_back_kernel(const double* __restrict__ in, long long ld_in, long long rows, int cols,
                                      double* __restrict__ out) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int c = c0 + y;
        const long long r = r0 + threadIdx.x;
        if (r < rows && c < cols) tile[y][threadIdx.x] = in[(long long)c * ld_in + r];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long r = r0 + y;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < cols) out[r * cols + c] = tile[threadIdx.x][y];
    }
}
cudaError_t launch_transpose_back(const double* in, long long ld_in, long long rows, int cols, double* out,
                                  cudaStream_t st, LaunchStats* ls) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
    transpose_back_kernel<<<grid, block, 0, st>>>(in, ld_in, rows, cols, out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// uniform [0,1) tangents from a counter-based generator keyed on the GLOBAL (row, site)
// index, so the field does not depend on how sites are sharded (pascal.random, fds/fds.py:23)
FDS_D uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void fill_random_kernel(double* __restrict__ T, long long n, int MO, int m, long long site_offset,
                                   long long n_global, uint64_t seed) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * MO) return;
    const long long i = idx / MO;
    const int j = (int)(idx % MO);
    if (j >= m) { T[idx] = 0.0; return; }
    const uint64_t ctr = (uint64_t)j * (uint64_t)n_global + (uint64_t)(site_offset + i);
    const uint64_t h = splitmix64(splitmix64(seed) ^ splitmix64(ctr));
    T[idx] = (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
cudaError_t launch_fill_random(double* T, long long n, int MO, int m, long long site_offset, long long n_global,
                               uint64_t seed, cudaStream_t st, LaunchStats* ls) {
    const long long total = n * MO;
    fill_random_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(T, n, MO, m, site_offset, n_global, seed);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// -------------------------------- Gram kernel -------------------------------- //
// 4x4 register blocks over the upper triangle of the (padded) MXP x MXP Gram.
// CTA = NG site-groups x GS lanes; lane l of a group owns block pair (bi[l], bj[l]);
// group g handles sites g, g+NG, ... of each staged tile.
constexpr int kGramTile = 32;   // sites staged per iteration

template <int SRC>
__global__ void __launch_bounds__(1024) gram_kernel(const double* __restrict__ E, const double* __restrict__ T,
                                                    const double* __restrict__ d, int m, double inv_eps, long long n,
                                                    int MXP, int nb, int GS, int NG, double* __restrict__ part,
                                                    BatchDesc bd) {
    extern __shared__ __align__(16) double sm[];      // Xs[kGramTile][MXP], then reduction scratch
    const int MX = m + 2, M = m + 2, MO = m + 1;
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        E += off * M; T += off * MO; d += off;
        part += (size_t)blockIdx.y * bd.part_stride;
    }
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int grp = tid / GS, lane = tid % GS;
    const int npair = nb * (nb + 1) / 2;
    int bi = 0, bj = 0;
    if (lane < npair) {                                // decode upper-triangular pair index
        int rem = lane;
        while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
        bj = bi + rem;
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const long long ntiles = (n + kGramTile - 1) / kGramTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * kGramTile;
        // stage X for kGramTile sites
        for (int e = tid; e < kGramTile * MXP; e += nthr) {
            const int s = e / MXP, l = e - s * MXP;
            const long long i = s0 + s;
            double v = 0.0;
            if (i < n) {
                if (l < MO) {
                    if (SRC == SRC_ENSEMBLE) v = (E[i * M + l + 1] - E[i * M]) * inv_eps;
                    else v = T[i * MO + l];
                } else if (l == MO) {
                    v = d[i];
                }
            }
            sm[e] = v;
        }
        __syncthreads();
        if (lane < npair && grp < NG) {
            for (int s = grp; s < kGramTile; s += NG) {
                const double* row = sm + s * MXP;
                const double2 a01 = *reinterpret_cast<const double2*>(row + 4 * bi);
                const double2 a23 = *reinterpret_cast<const double2*>(row + 4 * bi + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(row + 4 * bj);
                const double2 b23 = *reinterpret_cast<const double2*>(row + 4 * bj + 2);
                const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
            }
        }
        __syncthreads();
    }
    // fixed-order reduction over the site groups, then scatter into the full symmetric matrix
    double* red = sm;                                  // [NG][npair][16]
    if (lane < npair && grp < NG) {
        double* dst = red + ((size_t)grp * npair + lane) * 16;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dst[a * 4 + b] = acc[a][b];
    }
    __syncthreads();
    double* out = part + (size_t)blockIdx.x * MX * MX;
    for (int e = tid; e < npair * 16; e += nthr) {
        const int pr = e / 16, ab = e % 16, a = ab / 4, b = ab % 4;
        int pi = 0, rem = pr;
        while (rem >= nb - pi) { rem -= nb - pi; ++pi; }
        const int pj = pi + rem;
        double s = 0.0;
        for (int g = 0; g < NG; ++g) s += red[((size_t)g * npair + pr) * 16 + ab];
        const int row = 4 * pi + a, col = 4 * pj + b;
        if (row < MX && col < MX) {
            out[row * MX + col] = s;
            out[col * MX + row] = s;
        }
    }
}

__global__ void sum_partials_kernel(const double* __restrict__ part, int nparts, int len, double* __restrict__ out,
                                    BatchDesc bd) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    part += (size_t)blockIdx.y * bd.part_stride;
    out += (size_t)blockIdx.y * bd.small_stride;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    out[e] = s;
}

cudaError_t launch_gram(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                        long long n, double* part, int max_part, double* out, int num_sms, cudaStream_t st,
                        LaunchStats* ls, const BatchDesc* batch) {
    BatchDesc bd;
    if (batch) bd = *batch;
    const int MX = m + 2;
    const int MXP = (MX + 3) / 4 * 4;
    const int nb = MXP / 4;
    const int npair = nb * (nb + 1) / 2;
    const int GS = (npair + 31) / 32 * 32;
    if (GS > 1024) return cudaErrorInvalidValue;
    int NG = std::max(1, std::min(8, 512 / GS));
    const int block = GS * NG;
    const long long ntiles = (n + kGramTile - 1) / kGramTile;
    int grid = (int)std::min<long long>(ntiles, std::min(max_part, 2 * num_sms));
    if (grid < 1) grid = 1;
    const dim3 grid2((unsigned)grid, (unsigned)bd.count);
    const size_t smem = sizeof(double) * std::max((size_t)kGramTile * MXP, (size_t)NG * npair * 16);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(gram_kernel<SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_kernel<SRC_ENSEMBLE><<<grid2, block, smem, st>>>(E, T, d, m, inv_eps, n, MXP, nb, GS, NG, part, bd);
    } else {
        e = cudaFuncSetAttribute(gram_kernel<SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_kernel<SRC_TANGENT><<<grid2, block, smem, st>>>(E, T, d, m, inv_eps, n, MXP, nb, GS, NG, part, bd);
    }
    if (ls) ls->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int len = MX * MX;
    sum_partials_kernel<<<dim3((len + 255) / 256, bd.count), 256, 0, st>>>(part, grid, len, out, bd);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// -------------------------------- apply kernel ------------------------------- //
// out[i][o] = sum_l A[o][l] X_i[l];  4 sites x 4 outputs per thread, operands staged
// transposed in shared memory (Xt[l][S], At[l][MOP]) so both are read as 128-bit vectors.
constexpr int kApplyTile = 64;

template <int SRC>
__global__ void __launch_bounds__(512) apply_kernel(const double* __restrict__ E, const double* Tin,
                                                    const double* __restrict__ d, const double* __restrict__ A, int m,
                                                    double inv_eps, long long n, double* Tout, double* E_out,
                                                    const double* __restrict__ u, double eps, int MOP, BatchDesc bd) {
    extern __shared__ __align__(16) double sm[];
    const int MX = m + 2, M = m + 2, MO = m + 1;
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        E += off * M; Tin += off * MO; Tout += off * MO;
        if (d != nullptr) d += off;
        if (E_out != nullptr) { E_out += off * M; u += off; }
        A += (size_t)blockIdx.y * bd.small_stride;
    }
    double* At = sm;                         // [MX][MOP]
    double* Xt = sm + (size_t)MX * MOP;      // [MX][kApplyTile]
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < MX * MOP; e += nthr) {
        const int l = e / MOP, o = e - l * MOP;
        At[e] = o < MO ? A[o * MX + l] : 0.0;
    }
    const int nob = MOP / 4;                 // output blocks
    const int nsb = kApplyTile / 4;          // site blocks
    const int ob = tid % nob, sb = tid / nob;
    const long long ntiles = (n + kApplyTile - 1) / kApplyTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * kApplyTile;
        __syncthreads();
        for (int e = tid; e < kApplyTile * MX; e += nthr) {
            const int s = e / MX, l = e - s * MX;
            const long long i = s0 + s;
            double v = 0.0;
            if (i < n) {
                if (l < MO) {
                    if (SRC == SRC_ENSEMBLE) v = (E[i * M + l + 1] - E[i * M]) * inv_eps;
                    else v = Tin[i * MO + l];
                } else {
                    v = d != nullptr ? d[i] : 0.0;
                }
            }
            Xt[l * kApplyTile + s] = v;
        }
        __syncthreads();
        if (sb < nsb) {
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            for (int l = 0; l < MX; ++l) {
                const double2 x01 = *reinterpret_cast<const double2*>(Xt + l * kApplyTile + 4 * sb);
                const double2 x23 = *reinterpret_cast<const double2*>(Xt + l * kApplyTile + 4 * sb + 2);
                const double2 a01 = *reinterpret_cast<const double2*>(At + l * MOP + 4 * ob);
                const double2 a23 = *reinterpret_cast<const double2*>(At + l * MOP + 4 * ob + 2);
                const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
                const double av[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fma(xv[a], av[b], acc[a][b]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const long long i = s0 + 4 * sb + a;
                if (i >= n) continue;
                const double ui = E_out != nullptr ? u[i] : 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int o = 4 * ob + b;
                    if (o >= MO) continue;
                    Tout[i * MO + o] = acc[a][b];
                    if (E_out != nullptr) E_out[i * M + o + 1] = __dadd_rn(ui, __dmul_rn(acc[a][b], eps));
                }
                if (E_out != nullptr && ob == 0) E_out[i * M] = ui;
            }
        }
    }
}

cudaError_t launch_apply(int src, const double* E, const double* Tin, const double* d, const double* A, int m,
                         double inv_eps, long long n, double* Tout, double* E_out, const double* u, double eps,
                         cudaStream_t st, LaunchStats* ls, const BatchDesc* batch) {
    BatchDesc bd;
    if (batch) bd = *batch;
    const int MX = m + 2, MO = m + 1;
    const int MOP = (MO + 3) / 4 * 4;
    const int nob = MOP / 4, nsb = kApplyTile / 4;
    int block = (nob * nsb + 31) / 32 * 32;
    if (block > 512) return cudaErrorInvalidValue;
    const long long ntiles = (n + kApplyTile - 1) / kApplyTile;
    const dim3 grid((unsigned)std::min<long long>(ntiles, bd.count > 1 ? 8 : 148 * 8), (unsigned)bd.count);
    const size_t smem = sizeof(double) * ((size_t)MX * MOP + (size_t)MX * kApplyTile);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(apply_kernel<SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_kernel<SRC_ENSEMBLE><<<grid, block, smem, st>>>(E, Tin, d, A, m, inv_eps, n, Tout, E_out, u, eps, MOP, bd);
    } else {
        e = cudaFuncSetAttribute(apply_kernel<SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_kernel<SRC_TANGENT><<<grid, block, smem, st>>>(E, Tin, d, A, m, inv_eps, n, Tout, E_out, u, eps, MOP, bd);
    }
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// ------------------------------ small algebra (1 CTA) ------------------------ //
// In-place Cholesky G = L L^T of the leading mm x mm block of a matrix with row stride ld
// (lower triangle holds L on exit).  Returns via *fail the 1-based failing column, else 0.
// min_pivot: pivots at or below it count as failures (0: only non-positive ones).
__device__ void cta_cholesky(double* G, int ld, int mm, int* fail, double min_pivot = 0.0) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int j = 0; j < mm; ++j) {
        __syncthreads();
        const double piv = G[j * ld + j];
        if (!(piv > min_pivot)) { if (tid == 0 && *fail == 0) *fail = j + 1; }
        const double ljj = sqrt(piv > 0.0 ? piv : 1.0);
        __syncthreads();
        if (tid == 0) G[j * ld + j] = ljj;
        for (int i = j + 1 + tid; i < mm; i += nthr) G[i * ld + j] /= ljj;
        __syncthreads();
        const int rem = mm - j - 1;
        for (int e = tid; e < rem * rem; e += nthr) {
            const int i = j + 1 + e / rem, k = j + 1 + e % rem;
            if (k <= i) G[i * ld + k] -= G[i * ld + j] * G[k * ld + j];
        }
    }
    __syncthreads();
}

// Linv = L^{-1} for lower-triangular L (row stride ld); Linv is mm x mm, row stride mm, full storage
__device__ void cta_tri_inverse(const double* L, int ld, int mm, double* Linv) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < mm * mm; e += nthr) Linv[e] = 0.0;
    __syncthreads();
    for (int c = tid; c < mm; c += nthr) {          // column c of the inverse by forward substitution
        for (int i = c; i < mm; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= L[i * ld + k] * Linv[k * mm + c];
            Linv[i * mm + c] = s / L[i * ld + i];
        }
    }
    __syncthreads();
}

size_t small_smem_bytes(int m) {
    const int MX = m + 2;
    return sizeof(double) * ((size_t)MX * MX + 2 * (size_t)m * m + 4 * MX) + 16;
}

// SmallBufs of problem p of a batch
FDS_D SmallBufs small_of(SmallBufs B, long long stride, int p) {
    const size_t o = (size_t)p * stride;
    B.gram += o; B.A += o; B.L1 += o; B.R += o; B.b += o; B.gdil += o;
    B.status = reinterpret_cast<int*>(reinterpret_cast<double*>(B.status) + o);
    return B;
}

// Upper bound of cond_2 of the EQUILIBRATED Gram H = D G D, D = diag(G)^-1/2 (see small_single_kernel), in two parts
// around the in-place factorisation: |H|_2 <= min(|H|_inf, |H|_F) before, |H^-1|_2 likewise from L^-1 after.
// sq, rs, fr: m doubles of shared memory each.
__device__ double cta_eq_norm_before(const double* Gp, int m, double* sq, double* rs, double* fr, double* out) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int j = tid; j < m; j += nthr) sq[j] = sqrt(fmax(Gp[j * m + j], 0.0));
    __syncthreads();
    for (int j = tid; j < m; j += nthr) {
        double r = 0.0, f = 0.0;
        for (int l = 0; l < m; ++l) {
            const double sc = sq[j] * sq[l];
            const double h = sc > 0.0 ? Gp[j * m + l] / sc : (j == l ? 1.0 : 0.0);
            r += fabs(h);
            f += h * h;
        }
        rs[j] = r;
        fr[j] = f;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, f = 0.0;
        for (int j = 0; j < m; ++j) { a = fmax(a, rs[j]); f += fr[j]; }
        *out = fmin(a, sqrt(f));
    }
    __syncthreads();
    return *out;
}
__device__ double cta_eq_norm_after(const double* Li, int m, const double* sq, double* rs, double* fr, double* out) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int a = tid; a < m; a += nthr) {            // row sums of |H^-1|, H^-1 = D^-1 L^-T L^-1 D^-1
        double r = 0.0, f = 0.0;
        for (int b = 0; b < m; ++b) {
            double e = 0.0;
            for (int k = a > b ? a : b; k < m; ++k) e += Li[k * m + a] * Li[k * m + b];
            e *= sq[a] * sq[b];
            r += fabs(e);
            f += e * e;
        }
        rs[a] = r;
        fr[a] = f;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, f = 0.0;
        for (int j = 0; j < m; ++j) { a = fmax(a, rs[j]); f += fr[j]; }
        *out = fmin(a, sqrt(f));
    }
    __syncthreads();
    return *out;
}

// |S|_2 of a symmetric m x m matrix S (in A, row stride m; B: scratch of the same size; both are overwritten) from
// above: |S|_2^(2^k) = |S^(2^k)|_2 <= |S^(2^k)|_inf, so the 2^k-th root of the row-sum norm of S squared k times
// over-estimates |S|_2 by at most m^(1/2^(k+1)) (1.11 for m = 32, k = 4) -- against sqrt(m) for |S|_inf itself.
__device__ double cta_sym_norm2_bound(double* A, double* B, int m, int k, double* rs, double* out) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int it = 0; it < k; ++it) {
        for (int e = tid; e < m * m; e += nthr) {
            const int i = e / m, j = e % m;
            double acc = 0.0;
            for (int l = 0; l < m; ++l) acc += A[i * m + l] * A[l * m + j];      // (row i broadcast, row l at unit stride over the lanes)
            B[e] = acc;
        }
        __syncthreads();
        double* t = A; A = B; B = t;
    }
    for (int i = tid; i < m; i += nthr) {
        double r = 0.0;
        for (int j = 0; j < m; ++j) r += fabs(A[i * m + j]);
        rs[i] = r;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int i = 0; i < m; ++i) a = fmax(a, rs[i]);
        *out = pow(a, 1.0 / (double)(1 << k)) * (1.0 + 1e-9);       // (rounding of the k products: far below this margin)
    }
    __syncthreads();
    return *out;
}

// pass 1: G_dil/g_dil, projection, (optionally) first Cholesky factor; builds A1
// shift_rel > 0: the projected Gram gets shift_rel * trace added to its diagonal before it is factored (first pass of
// the shifted CholeskyQR3 below) -- the factor then exists for any numerically rank-deficient block
__global__ void __launch_bounds__(256) small1_kernel(SmallBufs B, int m, int use_d, int do_qr, long long bstride,
                                                     const int* gate, int gate_want, double shift_rel, long long gate_stride) {
    if (gate != nullptr && gate[(size_t)blockIdx.x * gate_stride] != gate_want) return;
    B = small_of(B, bstride, blockIdx.x);
    extern __shared__ __align__(16) double sm[];
    const int MX = m + 2, MO = m + 1, D = m + 1;
    double* G = sm;                      // MX*MX
    double* Gp = G + MX * MX;            // m*m
    double* Li = Gp + m * m;             // m*m
    double* c = Li + m * m;              // MX
    __shared__ int fail;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) fail = 0;
    for (int e = tid; e < MX * MX; e += nthr) G[e] = B.gram[e];
    __syncthreads();
    const double dd = G[D * MX + D];
    for (int k = tid; k < MO; k += nthr) {
        const double ck = use_d ? G[k * MX + D] / dd : 0.0;   // <t_k,d>/<d,d>: contribution AND projection coefficient
        c[k] = ck;
        B.gdil[k] = ck;
    }
    __syncthreads();
    if (!do_qr) {
        for (int e = tid; e < MO * MX; e += nthr) {
            const int o = e / MX, l = e % MX;
            B.A[e] = l == D ? -c[o] : (l == o ? 1.0 : 0.0);
        }
        if (tid == 0) *B.status = 0;
        return;
    }
    // Gram of the projected homogeneous tangents
    for (int e = tid; e < m * m; e += nthr) {
        const int j = e / m, l = e % m;
        Gp[e] = use_d ? G[j * MX + l] - c[j] * G[l * MX + D] : G[j * MX + l];
    }
    __syncthreads();
    if (shift_rel > 0.0) {
        __shared__ double shift;
        if (tid == 0) {
            double tr = 0.0;
            for (int j = 0; j < m; ++j) tr += Gp[j * m + j];
            shift = shift_rel * tr;
        }
        __syncthreads();
        for (int j = tid; j < m; j += nthr) Gp[j * m + j] += shift;
        __syncthreads();
    }
    // condition bound of the block as it is factored (equilibrated, see small_single_kernel): beyond ~1e7 the
    // caller must not trust CholeskyQR2 even when no pivot fails -- it switches to the shifted CholeskyQR3
    __shared__ double nrm1[2];
    double* rs = c + MX;
    double* sq = rs + MX;
    double* fr = sq + MX;
    cta_eq_norm_before(Gp, m, sq, rs, fr, &nrm1[0]);
    cta_cholesky(Gp, m, m, &fail);
    cta_tri_inverse(Gp, m, m, Li);
    cta_eq_norm_after(Li, m, sq, rs, fr, &nrm1[1]);
    if (tid == 0) reinterpret_cast<double*>(B.status)[1] = sqrt(nrm1[0] * nrm1[1]);
    for (int e = tid; e < m * m; e += nthr) {
        const int i = e / m, k = e % m;
        B.L1[e] = k <= i ? Gp[e] : 0.0;
    }
    for (int e = tid; e < MO * MX; e += nthr) {
        const int o = e / MX, l = e % MX;
        double v;
        if (o < m) {
            if (l < m) v = l <= o ? Li[o * m + l] : 0.0;
            else if (l == m) v = 0.0;
            else {                                    // d column: -(Linv c)_o
                double s = 0.0;
                for (int k = 0; k <= o; ++k) s += Li[o * m + k] * c[k];
                v = -s;
            }
        } else {                                      // inhomogeneous row: projection only
            v = l == m ? 1.0 : (l == D ? -c[m] : 0.0);
        }
        B.A[e] = v;
    }
    __syncthreads();
    if (tid == 0) *B.status = fail ? kSmallPass1 + fail : 0;
}

// Middle pass of the shifted CholeskyQR3 (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa 2020): the tangent array
// holds [Q1; p_w] with Q1 = L1s^-1 P from the SHIFTED first factor -- not orthonormal, but with cond_2 <~ 1e7 whatever
// cond_2(P) was.  Factor its Gram, Q2 = L2^-1 Q1 (inhomogeneous row untouched), and fold L2 into the accumulated
// factor: L1 <- L1 L2, so that the ordinary pass 2 (small2) ends with R = (L1s L2 L3)^T.
__global__ void __launch_bounds__(256) small_mid_kernel(SmallBufs B, int m, long long bstride) {
    B = small_of(B, bstride, blockIdx.x);
    extern __shared__ __align__(16) double sm[];
    const int MX = m + 2, MO = m + 1;
    double* G = sm;                      // MX*MX
    double* L2 = G + MX * MX;            // m*m
    double* Li = L2 + m * m;             // m*m
    __shared__ int fail;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) fail = 0;
    for (int e = tid; e < MX * MX; e += nthr) G[e] = B.gram[e];
    __syncthreads();
    for (int e = tid; e < m * m; e += nthr) L2[e] = G[(e / m) * MX + e % m];
    __syncthreads();
    // The Gram of Q1 has the eigenvalues lambda / (lambda + sigma) of the shifted first pass: tiny but well above
    // rounding for any block the method can handle, and (rounding aside) ZERO for an exactly rank-deficient block --
    // which must fail, not pass with a pivot made of rounding noise
    cta_cholesky(L2, m, m, &fail, 8.0 * m * 1.1102230246251565e-16);
    cta_tri_inverse(L2, m, m, Li);
    for (int e = tid; e < m * m; e += nthr) {        // (L1 L2)[i][k] = sum_{j = k .. i} L1[i][j] L2[j][k]   (G is free: scratch)
        const int i = e / m, k = e % m;
        double s = 0.0;
        if (k <= i)
            for (int j = k; j <= i; ++j) s += B.L1[i * m + j] * L2[j * m + k];
        G[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nthr) B.L1[e] = G[e];
    for (int e = tid; e < MO * MX; e += nthr) {
        const int o = e / MX, l = e % MX;
        double v = 0.0;
        if (o < m) { if (l < m && l <= o) v = Li[o * m + l]; }
        else if (l == m) v = 1.0;
        B.A[e] = v;
    }
    __syncthreads();
    if (tid == 0) *B.status = fail ? kSmallPass1 + fail : 0;
}

// pass 2: second Cholesky factor on the Gram of [Q1; p_w], b, R = (L1 L2)^T; builds A2
__global__ void __launch_bounds__(256) small2_kernel(SmallBufs B, int m, long long bstride, const double* gram2,
                                                     const int* gate, int gate_want, long long gate_stride) {
    if (gate != nullptr && gate[(size_t)blockIdx.x * gate_stride] != gate_want) return;
    if (gram2 != nullptr) gram2 += (size_t)blockIdx.x * bstride;
    B = small_of(B, bstride, blockIdx.x);
    if (gram2 != nullptr) B.gram = const_cast<double*>(gram2);
    extern __shared__ __align__(16) double sm[];
    const int MX = m + 2, MO = m + 1, D = m + 1;
    double* G = sm;                      // MX*MX
    double* L2 = G + MX * MX;            // m*m
    double* Li = L2 + m * m;             // m*m
    double* bb = Li + m * m;             // m
    __shared__ int fail;
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) fail = 0;
    for (int e = tid; e < MX * MX; e += nthr) G[e] = B.gram[e];
    __syncthreads();
    for (int e = tid; e < m * m; e += nthr) L2[e] = G[(e / m) * MX + e % m];
    __syncthreads();
    cta_cholesky(L2, m, m, &fail);
    cta_tri_inverse(L2, m, m, Li);
    for (int j = tid; j < m; j += nthr) {            // b = L2^{-1} (Q1 . p_w)
        double s = 0.0;
        for (int l = 0; l <= j; ++l) s += Li[j * m + l] * G[l * MX + m];
        bb[j] = s;
        B.b[j] = s;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nthr) {        // R[a][b] = sum_k L1[b][k] L2[k][a]  (upper triangular)
        const int a = e / m, b = e % m;
        double s = 0.0;
        if (b >= a)
            for (int k = a; k <= b; ++k) s += B.L1[b * m + k] * L2[k * m + a];
        B.R[e] = s;
    }
    for (int e = tid; e < MO * MX; e += nthr) {
        const int o = e / MX, l = e % MX;
        double v = 0.0;
        if (o < m) {
            if (l < m && l <= o) v = Li[o * m + l];
        } else {
            if (l < m) {                              // -(b^T L2^{-1})_l
                double s = 0.0;
                for (int j = l; j < m; ++j) s += bb[j] * Li[j * m + l];
                v = -s;
            } else if (l == m) v = 1.0;
        }
        B.A[e] = v;
    }
    (void)D;
    __syncthreads();
    if (tid == 0 && fail) *B.status = kSmallPass2 + fail;      // only ever raised: a pass-1 failure stays visible
}

// Single-pass CholeskyQR with its own certificate: everything the boundary needs -- G_dil, g_dil,
// the projection, Q = L^-1 P, R = L^T, b = Q p_w, w <- p_w - b^T Q -- from ONE extended Gram, as the
// constant matrix A.  The loss of orthogonality of one Cholesky pass is ~ cond_2(P)^2 u = cond_2(G) u
// with G = P P^T the projected Gram -- after equilibration (see below).  The kernel reports an upper bound of
// that condition number next to the status so that the two-pass CholeskyQR2 (small1 / small2) takes over
// whenever it is not negligible.
__global__ void __launch_bounds__(256) small_single_kernel(SmallBufs B, int m, int use_d, long long bstride,
                                                           int* two_pass_flag, int force_two, long long flag_stride) {
    if (two_pass_flag != nullptr) two_pass_flag += (size_t)blockIdx.x * flag_stride;
    B = small_of(B, bstride, blockIdx.x);
    extern __shared__ __align__(16) double sm[];
    const int MX = m + 2, MO = m + 1, D = m + 1;
    double* G = sm;                      // MX*MX
    double* Gp = G + MX * MX;            // m*m   projected Gram -> L
    double* Li = Gp + m * m;             // m*m
    double* c = Li + m * m;              // MX    projection coefficients
    double* bb = c + MX;                 // m     b, then b^T L^-1
    __shared__ int fail;
    __shared__ double nrm[2];
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) { fail = 0; nrm[0] = 0.0; nrm[1] = 0.0; }
    for (int e = tid; e < MX * MX; e += nthr) G[e] = B.gram[e];
    __syncthreads();
    const double dd = G[D * MX + D];
    for (int k = tid; k < MO; k += nthr) {
        const double ck = use_d ? G[k * MX + D] / dd : 0.0;
        c[k] = ck;
        B.gdil[k] = ck;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += nthr) {
        const int j = e / m, l = e % m;
        Gp[e] = use_d ? G[j * MX + l] - c[j] * G[l * MX + D] : G[j * MX + l];
    }
    __syncthreads();
    // Condition certificate on the EQUILIBRATED Gram H = D G D, D = diag(G)^-1/2 (unit diagonal): Cholesky's backward
    // error is componentwise relative to sqrt(G_ii G_jj), and Q = P L^-T does not change when the rows of P are
    // rescaled, so the orthogonality lost by one pass is governed by cond_2(H), not cond_2(G) (the rows of P have
    // grown by different factors over the segment: diag R spreads by 3 - 5).  |.|_2 <= min(|.|_inf, |.|_F) for the
    // symmetric H and H^-1 = D^-1 G^-1 D^-1 gives the bound; sq[j] = sqrt(G_jj).
    double* sq = bb + MX;                // m
    double* fr = sq + MX;                // m     squared Frobenius row sums
    cta_eq_norm_before(Gp, m, sq, bb, fr, &nrm[0]);
    cta_cholesky(Gp, m, m, &fail);
    cta_tri_inverse(Gp, m, m, Li);
    cta_eq_norm_after(Li, m, sq, bb, fr, &nrm[1]);
    // b = L^-1 <P_j, p_w>,  <P_j, p_w> = G[j][m] - c_j G[m][D]
    for (int j = tid; j < m; j += nthr) {
        double s = 0.0;
        for (int l = 0; l <= j; ++l) s += Li[j * m + l] * (use_d ? G[l * MX + m] - c[l] * G[m * MX + D] : G[l * MX + m]);
        B.b[j] = s;
        bb[j] = s;
    }
    for (int e = tid; e < m * m; e += nthr) {        // R = L^T (upper triangular); L1 kept for completeness
        const int a = e / m, b = e % m;
        B.R[e] = b >= a ? Gp[b * m + a] : 0.0;
        B.L1[e] = b <= a ? Gp[e] : 0.0;
    }
    __syncthreads();
    // q_l = (b^T L^-1)_l
    for (int e = tid; e < MO * MX; e += nthr) {
        const int o = e / MX, l = e % MX;
        double v = 0.0;
        if (o < m) {
            if (l < m) v = l <= o ? Li[o * m + l] : 0.0;
            else if (l == D) {                        // -(L^-1 c)_o
                double s = 0.0;
                for (int k = 0; k <= o; ++k) s += Li[o * m + k] * c[k];
                v = -s;
            }
        } else {
            if (l < m) {                              // -(b^T L^-1)_l
                double s = 0.0;
                for (int j = l; j < m; ++j) s += bb[j] * Li[j * m + l];
                v = -s;
            } else if (l == m) v = 1.0;
            else {                                    // d column: -c_m + sum_l (b^T L^-1)_l c_l
                double t = 0.0;
                for (int ll = 0; ll < m; ++ll) {
                    double s = 0.0;
                    for (int j = ll; j < m; ++j) s += bb[j] * Li[j * m + ll];
                    t += s * c[ll];
                }
                v = -c[m] + t;
            }
        }
        B.A[e] = v;
    }
    __syncthreads();
    // Tighten the certificate by repeated squaring (everything else has left shared memory: G and the factor's own
    // storage are scratch now): H = D L L^T D and H^-1 = D^-1 L^-T L^-1 D^-1, each squared four times.  The plain
    // row-sum / Frobenius bound over-estimates cond_2 by 2 - 4 at m = 32 and sent 2 of 30 well-conditioned
    // boundaries at N = 2^24 through CholeskyQR2.
    if (fail == 0) {
        __shared__ double nrm2[2];
        for (int e = tid; e < m * m; e += nthr) {
            const int j = e / m, l = e % m, kk = j < l ? j : l;
            double acc = 0.0;
            for (int k = 0; k <= kk; ++k) acc += Gp[j * m + k] * Gp[l * m + k];      // (lanes read a column of L: bank conflicts, but m^3 / 3 once)
            const double sc = sq[j] * sq[l];
            G[e] = sc > 0.0 ? acc / sc : (j == l ? 1.0 : 0.0);
        }
        __syncthreads();
        cta_sym_norm2_bound(G, Gp, m, 4, fr, &nrm2[0]);
        for (int e = tid; e < m * m; e += nthr) {
            const int a = e / m, b = e % m;
            double acc = 0.0;
            for (int k = a > b ? a : b; k < m; ++k) acc += Li[k * m + a] * Li[k * m + b];
            G[e] = acc * sq[a] * sq[b];
        }
        __syncthreads();
        cta_sym_norm2_bound(G, Gp, m, 4, fr, &nrm2[1]);
        if (tid == 0) {
            if (nrm2[0] < nrm[0]) nrm[0] = nrm2[0];
            if (nrm2[1] < nrm[1]) nrm[1] = nrm2[1];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double cond = sqrt(nrm[0] * nrm[1]);
        *B.status = fail;
        reinterpret_cast<double*>(B.status)[1] = cond;
        // one Cholesky pass loses orthogonality like cond^2 u: CholeskyQR2 takes over unless that is negligible
        if (two_pass_flag != nullptr) {
            const int two = (force_two || fail != 0 || !(cond * cond * 2.3e-16 < kSinglePassTol)) ? 1 : 0;
            if (flag_stride != 0) *two_pass_flag = two;          // this problem's own gate
            else if (two) atomicMax(two_pass_flag, 1);          // one gate for all (zeroed by the caller)
        }
    }
}

cudaError_t launch_small_single(const SmallBufs& B, int m, int use_d, cudaStream_t st, LaunchStats* ls,
                                const BatchDesc* batch, int* two_pass_flag, int force_two, long long flag_stride) {
    const size_t smem = small_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(small_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    small_single_kernel<<<batch ? batch->count : 1, 256, smem, st>>>(B, m, use_d, batch ? batch->small_stride : 0,
                                                                     two_pass_flag, force_two, flag_stride);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

cudaError_t launch_small1(const SmallBufs& B, int m, int use_d, int do_qr, cudaStream_t st, LaunchStats* ls,
                          const BatchDesc* batch, double shift_rel) {
    const size_t smem = small_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(small1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    small1_kernel<<<batch ? batch->count : 1, 256, smem, st>>>(B, m, use_d, do_qr, batch ? batch->small_stride : 0,
                                                               batch ? batch->gate : nullptr, batch ? batch->gate_want : 0,
                                                               shift_rel, batch ? batch->gate_stride : 0);
    if (ls) ls->launches++;
    return cudaGetLastError();
}
cudaError_t launch_small_mid(const SmallBufs& B, int m, cudaStream_t st, LaunchStats* ls, const BatchDesc* batch) {
    const size_t smem = small_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(small_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    small_mid_kernel<<<batch ? batch->count : 1, 256, smem, st>>>(B, m, batch ? batch->small_stride : 0);
    if (ls) ls->launches++;
    return cudaGetLastError();
}
cudaError_t launch_small2(const SmallBufs& B, int m, cudaStream_t st, LaunchStats* ls, const BatchDesc* batch,
                          const double* gram2) {
    const size_t smem = small_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(small2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    small2_kernel<<<batch ? batch->count : 1, 256, smem, st>>>(B, m, batch ? batch->small_stride : 0, gram2,
                                                               batch ? batch->gate : nullptr, batch ? batch->gate_want : 0,
                                                               batch ? batch->gate_stride : 0);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

}  // namespace fds
