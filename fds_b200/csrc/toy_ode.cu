// Small ODE systems of the reference test-suite as device solver plugins: Lorenz-63,
// Van der Pol, and the "circular" limit cycle (BASELINE configs 1-2; reference
// tests/solvers/lorenz/lorenz.f03:10-26, tests/solvers/vanderpol/vanderpol.f03:9-24,
// tests/solvers/circular/circular.f03:9-28; step-then-record loop
// tests/solvers/common/solver.f03:32-35).
//
// The state has 2 or 3 components, so there is nothing to parallelise inside one
// trajectory: thread k integrates ensemble column k for the whole chunk of steps in
// registers (one launch per segment instead of one per step) and records the objective
// after every step.  The ensemble keeps the site-major layout E[component][M] of the
// Lorenz-96 path, so the boundary kernels (Gram / CholeskyQR2 / apply) are shared.
//
// Operation order mirrors oracle/plugins.py (run_lorenz63, run_vanderpol, run_circular);
// ExactMath = no FMA contraction.  The Fortran DT literals are default REAL: float32
// values widened to double.
#include "kernels.h"

namespace fds {

template <class MP>
struct ToyLorenz63 {
    static constexpr int kDim = 3;
    static FDS_D void step(double* x, double s) {
        const double dt = (double)0.001f;
        const double s3 = 8. / 3;
        const double dx0 = MP::mul(10.0, MP::sub(x[1], x[0]));
        const double dx1 = MP::sub(MP::mul(x[0], MP::sub(s, x[2])), x[1]);
        const double dx2 = MP::sub(MP::mul(x[0], x[1]), MP::mul(s3, x[2]));
        x[0] = MP::add(x[0], MP::mul(dt, dx0));
        x[1] = MP::add(x[1], MP::mul(dt, dx1));
        x[2] = MP::add(x[2], MP::mul(dt, dx2));
    }
    // J = [(z - 28)^2, 100]   (reference tests/test_lorenz.py:18-31)
    static FDS_D void objective(const double* x, double& j0, double& j1) {
        const double d = MP::sub(x[2], 28.0);
        j0 = MP::mul(d, d);
        j1 = 100.0;
    }
};

template <class MP>
struct ToyVanDerPol {
    static constexpr int kDim = 2;
    static FDS_D void step(double* x, double s) {
        const double dt = (double)0.01f;
        const double dx0 = x[1];
        // (s/4) * (1 - x0^2) * x1 - x0
        const double a = MP::mul(MP::mul(s / 4, MP::sub(1.0, MP::mul(x[0], x[0]))), x[1]);
        const double dx1 = MP::sub(a, x[0]);
        x[0] = MP::add(x[0], MP::mul(dt, dx0));
        x[1] = MP::add(x[1], MP::mul(dt, dx1));
    }
    static FDS_D void objective(const double* x, double& j0, double& j1) {
        j0 = MP::mul(x[0], x[0]);
        j1 = 0.0;
    }
};

template <class MP>
struct ToyCircular {
    static constexpr int kDim = 2;
    static FDS_D void step(double* x, double s) {
        const double dt = (double)0.001f;
        const double r = MP::sub(MP::sub(MP::add(MP::mul(x[0], x[0]), MP::mul(x[1], x[1])), s), 1.0);
        const double dx0 = MP::sub(x[1], MP::mul(r, x[0]));
        const double dx1 = MP::sub(-x[0], MP::mul(r, x[1]));
        x[0] = MP::add(x[0], MP::mul(dt, dx0));
        x[1] = MP::add(x[1], MP::mul(dt, dx1));
    }
    static FDS_D void objective(const double* x, double& j0, double& j1) {
        j0 = MP::add(MP::mul(x[0], x[0]), MP::mul(x[1], x[1]));
        j1 = 0.0;
    }
};

// X, Y: [kDim][M] (may alias); column M-1 uses parameter s_last when M > 1 (the inhomogeneous
// trajectory of a segment, fds/segment.py:64).  jout: [nsteps][M][2] or nullptr.
template <class SYS>
__global__ void __launch_bounds__(128) toy_advance_kernel(const double* X, double* Y, int M, long long nsteps,
                                                          double s0, double s_last, double* __restrict__ jout) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const double s = (M > 1 && k == M - 1) ? s_last : s0;
    double x[SYS::kDim];
#pragma unroll
    for (int c = 0; c < SYS::kDim; ++c) x[c] = X[c * M + k];
    for (long long t = 0; t < nsteps; ++t) {
        SYS::step(x, s);
        if (jout != nullptr) {
            double j0, j1;
            SYS::objective(x, j0, j1);
            reinterpret_cast<double2*>(jout)[t * M + k] = make_double2(j0, j1);
        }
    }
#pragma unroll
    for (int c = 0; c < SYS::kDim; ++c) Y[c * M + k] = x[c];
}

int toy_dim(int system) {
    switch (system) {
        case 2: return 3;   // FDS_SYS_LORENZ63
        case 3: return 2;   // FDS_SYS_VANDERPOL
        case 4: return 2;   // FDS_SYS_CIRCULAR
    }
    return 0;
}

cudaError_t launch_toy_advance(int system, int exact, const double* X, double* Y, int M, long long nsteps,
                               double s0, double s_last, double* jout, cudaStream_t st, LaunchStats* ls) {
    if (nsteps < 0 || toy_dim(system) == 0) return cudaErrorInvalidValue;
    const int block = 128, grid = (M + block - 1) / block;
#define FDS_TOY(SYS) \
    do { if (exact) toy_advance_kernel<SYS<ExactMath>><<<grid, block, 0, st>>>(X, Y, M, nsteps, s0, s_last, jout); \
         else toy_advance_kernel<SYS<FastMath>><<<grid, block, 0, st>>>(X, Y, M, nsteps, s0, s_last, jout); } while (0)
    switch (system) {
        case 2: FDS_TOY(ToyLorenz63); break;
        case 3: FDS_TOY(ToyVanDerPol); break;
        case 4: FDS_TOY(ToyCircular); break;
    }
#undef FDS_TOY
    if (ls) ls->launches++;
    return cudaGetLastError();
}

}  // namespace fds
