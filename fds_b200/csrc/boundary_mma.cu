// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) versions of the two big passes of a segment
// boundary (K4 / K6 of SURVEY.md 2.3):
//
//   gram_mma    G = sum_i X_i X_i^T,  X_i = [t_0 .. t_m, d_i]  (MX = m+2 values per site).
//               The Gram IS a GEMM, X^T X with the sites as the contraction index: a warp takes
//               4 sites per instruction; lane (g, t) holds X[site t][8 cb + g] for every 8-column
//               block cb, and because A = X^T (row major) and B = X (column major) have the SAME
//               per-lane fragment, NBK loads feed NBK (NBK+1)/2 tensor-core instructions.
//               Shared-memory traffic per site drops from ~2.9 KB (4x4 register blocks) to the
//               320 B of the row itself, which is what makes the pass HBM-bound.
//
// Tiles arrive through the same TMA ring as boundary_fast.cu (tile_pipe.cuh) and are consumed
// straight from the raw stage: the finite-difference tangents (E_k - E_0) / eps are formed in
// registers on the way into the fragments.
//
// Reference: fds/timedilation.py:38-52 (contribution, project), fds/lsstan.py:20-34 (QR).
#include <algorithm>
#include <cstdlib>
#include "kernels.h"
#include "common.cuh"
#include "tile_pipe.cuh"

namespace fds {

FDS_D void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kMmaWarps = 8;          // consumer warps per CTA
constexpr int kMmaSites = 128;        // sites per tile (32 groups of 4)

// block = 8 consumer warps + 1 producer warp
template <int NBK, int SRC>
__global__ void __launch_bounds__(32 * (kMmaWarps + 1), 1) gram_mma_kernel(
    const double* __restrict__ E, const double* __restrict__ T, const double* __restrict__ d, int MX, double inv_eps,
    long long n, int nst, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NP = NBK * (NBK + 1) / 2, S = kMmaSites, NC = 32 * kMmaWarps;
    const int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    const int MO = MX - 1;
    TileStage ts;
    ts.full = reinterpret_cast<uint64_t*>(smem_raw);
    ts.empty = ts.full + kBndMaxStages;
    ts.raw = reinterpret_cast<double*>(smem_raw + 128);
    ts.stage_doubles = (S * W + S + 15) & ~15;
    ts.nst = nst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ts.nst; ++s) { mbar_init(&ts.full[s], 1); mbar_init(&ts.empty[s], kMmaWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (n + S - 1) / S;
    if (warp == kMmaWarps) { tile_producer<SRC>(ts, E, T, d, nullptr, MX, n, S, ntiles); return; }

    const int g = lane >> 2, t = lane & 3;
    // what column 8 cb + g of X is for this lane: 0 = tangent, 1 = d, 2 = padding (zero)
    int kind[NBK], off[NBK];
#pragma unroll
    for (int cb = 0; cb < NBK; ++cb) {
        const int c = 8 * cb + g;
        kind[cb] = c < MO ? 0 : (c == MO ? 1 : 2);
        off[cb] = SRC == SRC_ENSEMBLE ? c + 1 : c;
    }
    double acc[NP][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }

    int st = 0;
    uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int cnt = (int)((n - s0) < S ? (n - s0) : S);
        mbar_wait(&ts.full[st], par);
        const double* raw = ts.raw + (size_t)st * ts.stage_doubles;
#pragma unroll 2
        for (int grp = warp; grp < S / 4; grp += kMmaWarps) {
            const int site = 4 * grp + t;
            const bool valid = site < cnt;
            const double* row = raw + (size_t)site * W;
            const double e0 = SRC == SRC_ENSEMBLE ? row[0] : 0.0;
            const double dv = d != nullptr ? raw[(size_t)S * W + site] : 0.0;
            double x[NBK];
#pragma unroll
            for (int cb = 0; cb < NBK; ++cb) {
                double v = 0.0;
                if (kind[cb] == 0) v = SRC == SRC_ENSEMBLE ? (row[off[cb]] - e0) * inv_eps : row[off[cb]];
                else if (kind[cb] == 1) v = dv;
                x[cb] = valid ? v : 0.0;
            }
            int p = 0;
#pragma unroll
            for (int i = 0; i < NBK; ++i)
#pragma unroll
                for (int j = i; j < NBK; ++j, ++p) dmma_8x8x4(acc[p][0], acc[p][1], x[i], x[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);
        if (++st == ts.nst) { st = 0; par ^= 1u; }
    }
    // fixed-order reduction over the warps (the ring is drained: its memory is the scratch)
    named_barrier(1, NC);
    double* red = ts.raw;                         // [kMmaWarps][NP][64]
#pragma unroll
    for (int p = 0; p < NP; ++p)
        *reinterpret_cast<double2*>(red + ((size_t)warp * NP + p) * 64 + 2 * lane) = make_double2(acc[p][0], acc[p][1]);
    named_barrier(1, NC);
    double* out = part + (size_t)blockIdx.x * MX * MX;
    for (int e = tid; e < NP * 64; e += NC) {
        const int pr = e >> 6, l = (e & 63) >> 1, h = e & 1;
        int pi = 0, rem = pr;
        while (rem >= NBK - pi) { rem -= NBK - pi; ++pi; }
        const int pj = pi + rem;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) s += red[((size_t)w * NP + pr) * 64 + (e & 63)];
        // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1}
        const int r = 8 * pi + (l >> 2), c = 8 * pj + 2 * (l & 3) + h;
        if (r < MX && c < MX && r <= c) {
            out[r * MX + c] = s;
            out[c * MX + r] = s;
        }
    }
}

__global__ void sum_partials_mma_kernel(const double* __restrict__ part, int nparts, int len,
                                        double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    out[e] = s;
}

template <int NBK>
static cudaError_t gram_mma_launch(int src, const double* E, const double* T, const double* d, int MX,
                                   double inv_eps, long long n, double* part, int max_part, double* out,
                                   int num_sms, cudaStream_t st, LaunchStats* ls) {
    constexpr int NP = NBK * (NBK + 1) / 2, S = kMmaSites;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + S + 15) & ~(size_t)15;
    const size_t scratch = (size_t)kMmaWarps * NP * 64;
    int nst = (int)std::min<size_t>(kBndMaxStages, (200 * 1024) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = 128 + sizeof(double) * std::max(scratch, (size_t)nst * stage_doubles);
    const long long ntiles = (n + S - 1) / S;
    int grid = (int)std::min<long long>(ntiles, std::min(max_part, num_sms));
    if (grid < 1) grid = 1;
    const int block = 32 * (kMmaWarps + 1);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(gram_mma_kernel<NBK, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_mma_kernel<NBK, SRC_ENSEMBLE><<<grid, block, smem, st>>>(E, T, d, MX, inv_eps, n, nst, part);
    } else {
        e = cudaFuncSetAttribute(gram_mma_kernel<NBK, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_mma_kernel<NBK, SRC_TANGENT><<<grid, block, smem, st>>>(E, T, d, MX, inv_eps, n, nst, part);
    }
    if (ls) ls->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int len = MX * MX;
    sum_partials_mma_kernel<<<(len + 255) / 256, 256, 0, st>>>(part, grid, len, out);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

bool gram_mma_supported(int m) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("FDS_GRAM_MMA"); on = (e && e[0] == '0') ? 0 : 1; }
    return on && (m + 2 + 7) / 8 <= 7;
}

cudaError_t launch_gram_mma(int src, const double* E, const double* T, const double* d, int m, double inv_eps,
                            long long n, double* part, int max_part, double* out, int num_sms, cudaStream_t st,
                            LaunchStats* ls) {
    const int MX = m + 2;
#define FDS_GRAM_MMA_CASE(K) \
    case K: return gram_mma_launch<K>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls)
    switch ((MX + 7) / 8) {
        FDS_GRAM_MMA_CASE(1); FDS_GRAM_MMA_CASE(2); FDS_GRAM_MMA_CASE(3); FDS_GRAM_MMA_CASE(4);
        FDS_GRAM_MMA_CASE(5); FDS_GRAM_MMA_CASE(6); FDS_GRAM_MMA_CASE(7);
    }
#undef FDS_GRAM_MMA_CASE
    return cudaErrorInvalidValue;
}

}  // namespace fds
