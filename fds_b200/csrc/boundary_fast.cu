// Register-blocked FP64 kernels for the two big operations of a segment boundary
// (K4 / K6 of SURVEY.md 2.3), specialised at compile time on the row width
// MX = m + 2 of X_i = [t_0 .. t_m, d_i]:
//
//   gram_fast   G = sum_i X_i X_i^T.  Threads own B x B blocks of the upper triangle of G
//               (B = 5..7, so a block pair costs 2B shared-memory loads per B*B FMAs) and
//               sweep the sites of a staged tile; several sites are in flight per warp so
//               that (almost) every lane owns a block pair.
//   apply_fast  out_i = A X_i with the constant (m+1) x (m+2) matrix A passed BY VALUE as a
//               kernel parameter: one thread per site keeps X_i in registers and every FMA
//               takes its A operand straight from the constant bank -- no shared-memory
//               traffic in the inner loop.  Every A of a boundary is lower triangular in its
//               first m+1 columns plus the d column (projection, L^-1, b^T L^-1), which halves
//               the FMAs.  Rows are staged through shared memory so that global loads and
//               stores stay coalesced.
//
// Reference: fds/timedilation.py:38-52 (contribution, project), fds/lsstan.py:20-34 (QR, b,
// v update), fds/segment.py:43-51 (next perturbed initial conditions).
#include <algorithm>
#include <cstdlib>
#include "kernels.h"
#include "common.cuh"
#include "tile_pipe.cuh"

namespace fds {

// -------------------------------- Gram kernel -------------------------------- //
// block = 256 consumer threads + 1 producer warp
template <int B, int SRC>
__global__ void __launch_bounds__(288) gram_fast_kernel(const double* __restrict__ E, const double* __restrict__ T,
                                                        const double* __restrict__ d, int MX, double inv_eps,
                                                        long long n, int NB, int S, int nst, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NC = 256;
    const int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    const int MO = MX - 1;
    const int LD = NB * B + ((B & 1) ? 1 : 2);    // even B: rows and blocks 16-byte aligned -> 128-bit loads
    const int NP = NB * (NB + 1) / 2;             // block pairs of the upper triangle
    const int SLOTS = NC / NP;                    // sites in flight per iteration
    TileStage ts;
    ts.full = reinterpret_cast<uint64_t*>(smem_raw);
    ts.empty = ts.full + kBndMaxStages;
    double* xs = reinterpret_cast<double*>(smem_raw + 128);          // [S][LD] work buffer (also the reduction scratch)
    const size_t xs_doubles = (size_t)S * LD > (size_t)SLOTS * NP * B * B ? (size_t)S * LD : (size_t)SLOTS * NP * B * B;
    ts.raw = xs + ((xs_doubles + 15) & ~(size_t)15);
    ts.stage_doubles = (S * W + S + 15) & ~15;
    ts.nst = nst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ts.nst; ++s) { mbar_init(&ts.full[s], 1); mbar_init(&ts.empty[s], NC / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (n + S - 1) / S;
    if (warp == NC / 32) { tile_producer<SRC>(ts, E, T, d, nullptr, MX, n, S, ntiles); return; }

    const int slot = tid / NP, pair = tid - slot * NP;
    int bi = 0, bj = 0;
    {
        int rem = pair;
        while (rem >= NB - bi) { rem -= NB - bi; ++bi; }
        bj = bi + rem;
    }
    const bool worker = slot < SLOTS;
    double acc[B][B];
#pragma unroll
    for (int a = 0; a < B; ++a)
#pragma unroll
        for (int b = 0; b < B; ++b) acc[a][b] = 0.0;

    int st = 0;
    uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int cnt = (int)((n - s0) < S ? (n - s0) : S);
        mbar_wait(&ts.full[st], par);
        const double* raw = ts.raw + (size_t)st * ts.stage_doubles;
        // raw rows -> X rows [t_0 .. t_m, d, 0 ..] in the work buffer: thread (ty, tx) converts column
        // tx of rows ty, ty+4, ...; four rows in flight
        {
            const double* __restrict__ rw = raw;
            double* __restrict__ xw = xs;
            const int tx = tid & 63, ty = tid >> 6;
            if (tx < LD) {
#pragma unroll 4
                for (int s = ty; s < S; s += NC / 64) {
                    const double* rr = rw + (size_t)s * W;
                    double v = 0.0;
                    if (s < cnt) {
                        if (tx < MO) v = SRC == SRC_ENSEMBLE ? (rr[tx + 1] - rr[0]) * inv_eps : rr[tx];
                        else if (tx == MO) v = d != nullptr ? rw[(size_t)S * W + s] : 0.0;
                    }
                    xw[s * LD + tx] = v;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);              // the raw stage may be refilled
        if (++st == ts.nst) { st = 0; par ^= 1u; }
        named_barrier(1, NC);
        if (worker) {
#pragma unroll 2
            for (int s = slot; s < S; s += SLOTS) {
                const double* row = xs + s * LD;
                double av[B], bv[B];
                if ((B & 1) == 0) {
#pragma unroll
                    for (int a = 0; a < B; a += 2) {
                        const double2 pa = *reinterpret_cast<const double2*>(row + bi * B + a);
                        const double2 pb = *reinterpret_cast<const double2*>(row + bj * B + a);
                        av[a] = pa.x; av[a + 1] = pa.y; bv[a] = pb.x; bv[a + 1] = pb.y;
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < B; ++a) { av[a] = row[bi * B + a]; bv[a] = row[bj * B + a]; }
                }
#pragma unroll
                for (int a = 0; a < B; ++a)
#pragma unroll
                    for (int b = 0; b < B; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
            }
        }
        named_barrier(1, NC);
    }
    // fixed-order reduction over the slots, then scatter into the full symmetric matrix
    double* red = xs;                             // [SLOTS][NP][B*B]
    if (worker) {
        double* dst = red + ((size_t)slot * NP + pair) * (B * B);
#pragma unroll
        for (int a = 0; a < B; ++a)
#pragma unroll
            for (int b = 0; b < B; ++b) dst[a * B + b] = acc[a][b];
    }
    named_barrier(1, NC);
    double* out = part + (size_t)blockIdx.x * MX * MX;
    for (int e = tid; e < NP * B * B; e += NC) {
        const int pr = e / (B * B), ab = e - pr * (B * B), a = ab / B, b = ab - a * B;
        int pi = 0, rem = pr;
        while (rem >= NB - pi) { rem -= NB - pi; ++pi; }
        const int pj = pi + rem;
        double s = 0.0;
        for (int g = 0; g < SLOTS; ++g) s += red[((size_t)g * NP + pr) * (B * B) + ab];
        const int r = pi * B + a, c = pj * B + b;
        if (r < MX && c < MX) {
            out[r * MX + c] = s;
            out[c * MX + r] = s;
        }
    }
}

__global__ void sum_partials_fast_kernel(const double* __restrict__ part, int nparts, int len,
                                         double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    out[e] = s;
}

template <int B>
static cudaError_t gram_fast_launch(int src, const double* E, const double* T, const double* d, int MX,
                                    double inv_eps, long long n, double* part, int max_part, double* out,
                                    int num_sms, cudaStream_t st, LaunchStats* ls) {
    const int NB = (MX + B - 1) / B;
    const int NP = NB * (NB + 1) / 2;
    const int LD = NB * B + ((B & 1) ? 1 : 2);
    const int SLOTS = 256 / NP;
    if (SLOTS < 1) return cudaErrorInvalidValue;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    int S = std::min(SLOTS * 8, 128);             // about 8 sites per slot and tile
    S = (S + 1) & ~1;                             // even: every tile starts 16-byte aligned
    const size_t xs_doubles = std::max((size_t)S * LD, (size_t)SLOTS * NP * B * B);
    const size_t stage_doubles = ((size_t)S * W + S + 15) & ~(size_t)15;
    const size_t fixed = 128 + sizeof(double) * ((xs_doubles + 15) & ~(size_t)15);
    const int nst = (int)std::max<size_t>(2, std::min<size_t>(kBndMaxStages, (100 * 1024 - std::min<size_t>(fixed, 100 * 1024)) / (stage_doubles * sizeof(double))));
    const size_t smem = fixed + sizeof(double) * nst * stage_doubles;
    const long long ntiles = (n + S - 1) / S;
    cudaError_t e;
    int occ = 1;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(gram_fast_kernel<B, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gram_fast_kernel<B, SRC_ENSEMBLE>, 288, smem);
    } else {
        e = cudaFuncSetAttribute(gram_fast_kernel<B, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gram_fast_kernel<B, SRC_TANGENT>, 288, smem);
    }
    int grid = (int)std::min<long long>(ntiles, std::min(max_part, std::max(1, occ) * num_sms));   // one resident wave
    if (grid < 1) grid = 1;
    if (src == SRC_ENSEMBLE)
        gram_fast_kernel<B, SRC_ENSEMBLE><<<grid, 288, smem, st>>>(E, T, d, MX, inv_eps, n, NB, S, nst, part);
    else
        gram_fast_kernel<B, SRC_TANGENT><<<grid, 288, smem, st>>>(E, T, d, MX, inv_eps, n, NB, S, nst, part);
    if (ls) ls->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int len = MX * MX;
    sum_partials_fast_kernel<<<(len + 255) / 256, 256, 0, st>>>(part, grid, len, out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

// Block size of the register tile.  Measured on B200 at m = 32 (N = 2^24): B = 4 (72 registers,
// several CTAs per SM, 128-bit shared loads) 19.2 ms per boundary, B = 5: 19.6, B = 6..8: 23 --
// occupancy beats the lower load-per-FMA ratio of the larger tiles.  FDS_GRAM_B overrides.
static int pick_gram_block(int MX) {
    if (const char* e = getenv("FDS_GRAM_B")) { const int b = atoi(e); if (b >= 4 && b <= 8) return b; }
    (void)MX;
    return 4;
}

cudaError_t launch_gram_fast(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                             long long n, double* part, int max_part, double* out, int num_sms, cudaStream_t st,
                             LaunchStats* ls) {
    const int MX = m + 2;
    switch (pick_gram_block(MX)) {
        case 4: return gram_fast_launch<4>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls);
        case 5: return gram_fast_launch<5>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls);
        case 6: return gram_fast_launch<6>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls);
        case 7: return gram_fast_launch<7>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls);
        case 8: return gram_fast_launch<8>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls);
    }
    return cudaErrorInvalidValue;
}

// -------------------------------- apply kernel ------------------------------- //
template <int MX>
struct ApplyMat { double a[(MX - 1) * MX]; };     // (m+1) x (m+2), row major

constexpr int kApplySites = 256;                  // sites per tile = consumer threads per CTA

// block = 256 consumer threads (one per site of a tile) + 1 producer warp
template <int MX, int SRC>
__global__ void __launch_bounds__(kApplySites + 32, 1) apply_fast_kernel(
    const double* __restrict__ E, const double* Tin, const double* __restrict__ d,
    const __grid_constant__ ApplyMat<MX> A, double inv_eps, long long n, double* Tout, double* E_out,
    const double* __restrict__ u, double eps, int nst) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int MO = MX - 1, M = MX, D = MX - 1, NC = kApplySites, S = kApplySites;
    constexpr int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    constexpr int LD = MX | 1;                    // odd row stride of the output buffer
    TileStage ts;
    ts.full = reinterpret_cast<uint64_t*>(smem_raw);
    ts.empty = ts.full + kBndMaxStages;
    double* os = reinterpret_cast<double*>(smem_raw + 128);          // [S][LD] output rows
    ts.raw = os + (((size_t)S * LD + 15) & ~(size_t)15);
    ts.stage_doubles = (S * W + 2 * S + 15) & ~15;
    ts.nst = nst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ts.nst; ++s) { mbar_init(&ts.full[s], 1); mbar_init(&ts.empty[s], NC / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (n + S - 1) / S;
    if (warp == NC / 32) { tile_producer<SRC>(ts, E, Tin, d, E_out != nullptr ? u : nullptr, MX, n, S, ntiles); return; }

    int st = 0;
    uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int cnt = (int)((n - s0) < S ? (n - s0) : S);
        mbar_wait(&ts.full[st], par);
        const double* raw = ts.raw + (size_t)st * ts.stage_doubles;
        // one thread per site: X_i into registers
        double x[MX];
        {
            const int s = tid < cnt ? tid : 0;    // threads past the end recompute site 0 of the tile, unstored
            const double* rr = raw + (size_t)s * W;
            if (SRC == SRC_ENSEMBLE) {
                const double e0 = rr[0];
#pragma unroll
                for (int l = 0; l < MO; ++l) x[l] = (rr[l + 1] - e0) * inv_eps;
            } else {
#pragma unroll
                for (int l = 0; l < MO; ++l) x[l] = rr[l];
            }
            x[D] = d != nullptr ? raw[(size_t)S * W + s] : 0.0;
            if (E_out != nullptr) os[tid * LD + MO] = raw[(size_t)S * W + S + s];    // u_i rides in the spare column
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);              // the raw stage may be refilled
        if (++st == ts.nst) { st = 0; par ^= 1u; }
        // out_i = A X_i, A lower triangular in its first MO columns + the d column
        double* row = os + tid * LD;
#pragma unroll
        for (int o = 0; o < MO; ++o) {
            double acc = A.a[o * MX + D] * x[D];
#pragma unroll
            for (int l = 0; l <= o; ++l) acc = fma(A.a[o * MX + l], x[l], acc);
            row[o] = acc;
        }
        named_barrier(1, NC);
        // coalesced stores: explicit tangents and / or the next perturbed ensemble
        const int tx = tid & 63, ty = tid >> 6;
        for (int s = ty; s < cnt; s += NC / 64) {
            const long long i = s0 + s;
            if (Tout != nullptr && tx < MO) Tout[i * MO + tx] = os[s * LD + tx];
            if (E_out != nullptr && tx < M) {
                const double ui = os[s * LD + MO];
                E_out[i * M + tx] = tx == 0 ? ui : __dadd_rn(ui, __dmul_rn(os[s * LD + tx - 1], eps));
            }
        }
        named_barrier(1, NC);
    }
}

template <int MX>
static cudaError_t apply_fast_launch(int src, const double* E, const double* Tin, const double* d,
                                     const double* A_host, double inv_eps, long long n, double* Tout,
                                     double* E_out, const double* u, double eps, int num_sms, cudaStream_t st,
                                     LaunchStats* ls) {
    ApplyMat<MX> A;
    for (int e = 0; e < (MX - 1) * MX; ++e) A.a[e] = A_host[e];
    constexpr int LD = MX | 1, S = kApplySites;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + 2 * S + 15) & ~(size_t)15;
    const size_t fixed = 128 + sizeof(double) * (((size_t)S * LD + 15) & ~(size_t)15);
    const int nst = (int)std::min<size_t>(kBndMaxStages, (226 * 1024 - fixed) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = fixed + sizeof(double) * nst * stage_doubles;
    const long long ntiles = (n + S - 1) / S;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(apply_fast_kernel<MX, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_fast_kernel<MX, SRC_ENSEMBLE><<<grid, S + 32, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst);
    } else {
        e = cudaFuncSetAttribute(apply_fast_kernel<MX, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_fast_kernel<MX, SRC_TANGENT><<<grid, S + 32, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst);
    }
    if (ls) ls->launches++;
    return cudaGetLastError();
}

bool apply_fast_supported(int m) {
    // (m + 2 = 50 would need two 106-KB stages: left to the generic kernels)
    switch (m + 2) { case 3: case 4: case 5: case 6: case 7: case 10: case 18: case 34: return true; }
    return false;
}

// A_host: the (m+1) x (m+2) matrix on the HOST (it travels as a kernel parameter)
cudaError_t launch_apply_fast(int src, const double* E, const double* Tin, const double* d, const double* A_host,
                              int m, double inv_eps, long long n, double* Tout, double* E_out, const double* u,
                              double eps, int num_sms, cudaStream_t st, LaunchStats* ls) {
#define FDS_APPLY_CASE(MXV) \
    case MXV: return apply_fast_launch<MXV>(src, E, Tin, d, A_host, inv_eps, n, Tout, E_out, u, eps, num_sms, st, ls)
    switch (m + 2) {
        FDS_APPLY_CASE(3); FDS_APPLY_CASE(4); FDS_APPLY_CASE(5); FDS_APPLY_CASE(6); FDS_APPLY_CASE(7);
        FDS_APPLY_CASE(10); FDS_APPLY_CASE(18); FDS_APPLY_CASE(34);
    }
#undef FDS_APPLY_CASE
    return cudaErrorInvalidValue;
}

}  // namespace fds
