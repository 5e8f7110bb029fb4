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
    long long n, int nst, double* __restrict__ part, BatchDesc bd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (bd.gate != nullptr && bd.gate[(size_t)blockIdx.y * bd.gate_stride] != bd.gate_want) return;      // the untaken CholeskyQR variant
    constexpr int NP = NBK * (NBK + 1) / 2, S = kMmaSites, NC = 32 * kMmaWarps;
    const int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    const int MO = MX - 1;
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        if (SRC == SRC_ENSEMBLE) E += off * MX; else T += off * MO;
        if (d != nullptr) d += off;
        part += (size_t)blockIdx.y * bd.part_stride;
    }
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


// The same Gram with the LAST NV (1 .. 4) columns of X taken off the tensor cores.  DMMA runs at the FMA rate of the
// vector FP64 pipe on this part, so a padded 8-column block costs what it would cost full: at MX = 34 the fifth
// block holds 2 live columns (the inhomogeneous tangent and d) and takes 5 of the 15 instructions of a group -- the
// pass is then bound by the FP64 pipe at 1.7x its HBM time.  Here the NBD = MX / 8 full blocks go through
// NBD (NBD + 1) / 2 DMMA as before and the NV leftover columns through plain DFMA: lane (g, t) already holds
// X[site t][8 cb + g], so  va[e][cb] += ex_e[site t] * x[cb]  is one FMA per (e, cb) and lane; the sums over the
// four sites of a group (lanes t = 0 .. 3) are folded by two shuffles at the very end.  11 DFMA (352 FMA slots)
// replace 5 DMMA (1280 FMA slots) per group of 4 sites at MX = 34.
template <int NBD, int NV, int SRC>
__global__ void __launch_bounds__(32 * (kMmaWarps + 1), 2) gram_mma2_kernel(
    const double* __restrict__ E, const double* __restrict__ T, const double* __restrict__ d, int MX, double inv_eps,
    long long n, int nst, double* __restrict__ part, BatchDesc bd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (bd.gate != nullptr && bd.gate[(size_t)blockIdx.y * bd.gate_stride] != bd.gate_want) return;      // the untaken CholeskyQR variant
    constexpr int NP = NBD * (NBD + 1) / 2, S = kMmaSites, NC = 32 * kMmaWarps, NS = NV * (NV + 1) / 2;
    const int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    const int MO = MX - 1;
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        if (SRC == SRC_ENSEMBLE) E += off * MX; else T += off * MO;
        if (d != nullptr) d += off;
        part += (size_t)blockIdx.y * bd.part_stride;
    }
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
    const int off0 = (SRC == SRC_ENSEMBLE ? 1 : 0) + g;       // raw-row offset of column 8 cb + g: off0 + 8 cb
    double acc[NP][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }
    double va[NV][NBD], vs[NS];
#pragma unroll
    for (int e = 0; e < NV; ++e)
#pragma unroll
        for (int cb = 0; cb < NBD; ++cb) va[e][cb] = 0.0;
#pragma unroll
    for (int q = 0; q < NS; ++q) vs[q] = 0.0;

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
            double x[NBD], ex[NV];
#pragma unroll
            for (int cb = 0; cb < NBD; ++cb) {
                const double v = SRC == SRC_ENSEMBLE ? (row[off0 + 8 * cb] - e0) * inv_eps : row[off0 + 8 * cb];
                x[cb] = valid ? v : 0.0;
            }
#pragma unroll
            for (int e = 0; e < NV; ++e) {
                const int c = 8 * NBD + e;                    // tangent column, or the d column (c == MO)
                double v;
                if (c < MO) v = SRC == SRC_ENSEMBLE ? (row[c + 1] - e0) * inv_eps : row[c];
                else v = d != nullptr ? raw[(size_t)S * W + site] : 0.0;
                ex[e] = valid ? v : 0.0;
            }
            int p = 0;
#pragma unroll
            for (int i = 0; i < NBD; ++i)
#pragma unroll
                for (int j = i; j < NBD; ++j, ++p) dmma_8x8x4(acc[p][0], acc[p][1], x[i], x[j]);
#pragma unroll
            for (int e = 0; e < NV; ++e)
#pragma unroll
                for (int cb = 0; cb < NBD; ++cb) va[e][cb] = fma(ex[e], x[cb], va[e][cb]);
            int q = 0;
#pragma unroll
            for (int e = 0; e < NV; ++e)
#pragma unroll
                for (int f = e; f < NV; ++f, ++q) vs[q] = fma(ex[e], ex[f], vs[q]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);
        if (++st == ts.nst) { st = 0; par ^= 1u; }
    }
    // the four sites of a group live in lanes t = 0 .. 3: fold them (fixed order)
#pragma unroll
    for (int e = 0; e < NV; ++e)
#pragma unroll
        for (int cb = 0; cb < NBD; ++cb) {
            va[e][cb] += __shfl_xor_sync(0xffffffffu, va[e][cb], 1);
            va[e][cb] += __shfl_xor_sync(0xffffffffu, va[e][cb], 2);
        }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        vs[q] += __shfl_xor_sync(0xffffffffu, vs[q], 1);
        vs[q] += __shfl_xor_sync(0xffffffffu, vs[q], 2);
    }
    // fixed-order reduction over the warps (the ring is drained: its memory is the scratch)
    named_barrier(1, NC);
    double* red = ts.raw;                                     // [kMmaWarps][NP][64]
    double* vred = red + (size_t)kMmaWarps * NP * 64;         // [kMmaWarps][NV][8 NBD]
    double* sred = vred + (size_t)kMmaWarps * NV * 8 * NBD;   // [kMmaWarps][NS]
#pragma unroll
    for (int p = 0; p < NP; ++p)
        *reinterpret_cast<double2*>(red + ((size_t)warp * NP + p) * 64 + 2 * lane) = make_double2(acc[p][0], acc[p][1]);
    if (t == 0) {
#pragma unroll
        for (int e = 0; e < NV; ++e)
#pragma unroll
            for (int cb = 0; cb < NBD; ++cb) vred[((size_t)warp * NV + e) * 8 * NBD + 8 * cb + g] = va[e][cb];
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) sred[warp * NS + q] = vs[q];
    }
    named_barrier(1, NC);
    double* out = part + (size_t)blockIdx.x * MX * MX;
    for (int e = tid; e < NP * 64; e += NC) {
        const int pr = e >> 6, l = (e & 63) >> 1, h = e & 1;
        int pi = 0, rem = pr;
        while (rem >= NBD - pi) { rem -= NBD - pi; ++pi; }
        const int pj = pi + rem;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) s += red[((size_t)w * NP + pr) * 64 + (e & 63)];
        // C fragment: row = lane / 4, columns 2 (lane % 4) + {0, 1}
        const int r = 8 * pi + (l >> 2), c = 8 * pj + 2 * (l & 3) + h;
        if (r <= c) {
            out[r * MX + c] = s;
            out[c * MX + r] = s;
        }
    }
    for (int e = tid; e < NV * 8 * NBD; e += NC) {
        const int ev = e / (8 * NBD), col = e - ev * 8 * NBD;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) s += vred[((size_t)w * NV + ev) * 8 * NBD + col];
        const int r = 8 * NBD + ev;
        out[r * MX + col] = s;
        out[col * MX + r] = s;
    }
    if (tid < NS) {
        int ev = 0, rem = tid;
        while (rem >= NV - ev) { rem -= NV - ev; ++ev; }
        const int fv = ev + rem;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) s += sred[w * NS + tid];
        out[(8 * NBD + ev) * MX + 8 * NBD + fv] = s;
        out[(8 * NBD + fv) * MX + 8 * NBD + ev] = s;
    }
}

__global__ void sum_partials_mma_kernel(const double* __restrict__ part, int nparts, int len,
                                        double* __restrict__ out, BatchDesc bd) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    if (bd.gate != nullptr && bd.gate[(size_t)blockIdx.y * bd.gate_stride] != bd.gate_want) return;
    part += (size_t)blockIdx.y * bd.part_stride;
    out += (size_t)blockIdx.y * bd.small_stride;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * len + e];
    out[e] = s;
}

static bool gram_split_on() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("FDS_GRAM_SPLIT"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

template <int NBD, int NV, int SRC>
static cudaError_t gram_mma2_launch_src(const double* E, const double* T, const double* d, int MX, double inv_eps,
                                        long long n, int nst, double* part, dim3 grid, int block, size_t smem,
                                        cudaStream_t st, const BatchDesc& bd) {
    cudaError_t e = cudaFuncSetAttribute(gram_mma2_kernel<NBD, NV, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gram_mma2_kernel<NBD, NV, SRC><<<grid, block, smem, st>>>(E, T, d, MX, inv_eps, n, nst, part, bd);
    return cudaGetLastError();
}

// MX = 8 NBD + NV, 1 <= NV <= 4: the leftover columns on the vector pipe (gram_mma2_kernel)
template <int NBD, int NV>
static cudaError_t gram_mma2_launch(int src, const double* E, const double* T, const double* d, int MX,
                                    double inv_eps, long long n, double* part, int max_part, double* out,
                                    int num_sms, cudaStream_t st, LaunchStats* ls, const BatchDesc& bd) {
    constexpr int NP = NBD * (NBD + 1) / 2, S = kMmaSites, NS = NV * (NV + 1) / 2;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + S + 15) & ~(size_t)15;
    const size_t scratch = (size_t)kMmaWarps * (NP * 64 + NV * 8 * NBD + NS);
    // two CTAs per SM (96 registers, half the ring each): one CTA's tensor-core work overlaps the other's waits --
    // with one CTA the pass sat at 57 % of the DMMA rate AND 56 % of the DRAM rate (profiles/r2b_boundary_ncu.txt)
    static const int ctas = [] { const char* e = getenv("FDS_GRAM_CTAS"); return (e && e[0] == '1') ? 1 : 2; }();
    int nst = (int)std::min<size_t>(kBndMaxStages, ((ctas == 2 ? 104 : 200) * 1024) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = 128 + sizeof(double) * std::max(scratch, (size_t)nst * stage_doubles);
    const long long ntiles = (n + S - 1) / S;
    int gx = (int)std::min<long long>(ntiles, std::min(max_part, ctas * num_sms));
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)bd.count);
    const int block = 32 * (kMmaWarps + 1);
    cudaError_t e = src == SRC_ENSEMBLE
        ? gram_mma2_launch_src<NBD, NV, SRC_ENSEMBLE>(E, T, d, MX, inv_eps, n, nst, part, grid, block, smem, st, bd)
        : gram_mma2_launch_src<NBD, NV, SRC_TANGENT>(E, T, d, MX, inv_eps, n, nst, part, grid, block, smem, st, bd);
    if (ls) ls->launches++;
    if (e != cudaSuccess) return e;
    const int len = MX * MX;
    sum_partials_mma_kernel<<<dim3((len + 255) / 256, bd.count), 256, 0, st>>>(part, gx, len, out, bd);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

template <int NBK>
static cudaError_t gram_mma_launch(int src, const double* E, const double* T, const double* d, int MX,
                                   double inv_eps, long long n, double* part, int max_part, double* out,
                                   int num_sms, cudaStream_t st, LaunchStats* ls, const BatchDesc& bd) {
    constexpr int NP = NBK * (NBK + 1) / 2, S = kMmaSites;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + S + 15) & ~(size_t)15;
    const size_t scratch = (size_t)kMmaWarps * NP * 64;
    int nst = (int)std::min<size_t>(kBndMaxStages, (200 * 1024) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = 128 + sizeof(double) * std::max(scratch, (size_t)nst * stage_doubles);
    const long long ntiles = (n + S - 1) / S;
    int gx = (int)std::min<long long>(ntiles, std::min(max_part, num_sms));
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, (unsigned)bd.count);
    const int block = 32 * (kMmaWarps + 1);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(gram_mma_kernel<NBK, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_mma_kernel<NBK, SRC_ENSEMBLE><<<grid, block, smem, st>>>(E, T, d, MX, inv_eps, n, nst, part, bd);
    } else {
        e = cudaFuncSetAttribute(gram_mma_kernel<NBK, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gram_mma_kernel<NBK, SRC_TANGENT><<<grid, block, smem, st>>>(E, T, d, MX, inv_eps, n, nst, part, bd);
    }
    if (ls) ls->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int len = MX * MX;
    sum_partials_mma_kernel<<<dim3((len + 255) / 256, bd.count), 256, 0, st>>>(part, gx, len, out, bd);
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
                            LaunchStats* ls, const BatchDesc* batch) {
    const int MX = m + 2;
    BatchDesc bd;
    if (batch) bd = *batch;
    if (gram_split_on()) {
        // the two production shapes: m = 32 (34 = 4 x 8 + 2) and m = 16 (18 = 2 x 8 + 2)
        if (MX == 34) return gram_mma2_launch<4, 2>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls, bd);
        if (MX == 18) return gram_mma2_launch<2, 2>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls, bd);
    }
#define FDS_GRAM_MMA_CASE(K) \
    case K: return gram_mma_launch<K>(src, E, T, d, MX, inv_eps, n, part, max_part, out, num_sms, st, ls, bd)
    switch ((MX + 7) / 8) {
        FDS_GRAM_MMA_CASE(1); FDS_GRAM_MMA_CASE(2); FDS_GRAM_MMA_CASE(3); FDS_GRAM_MMA_CASE(4);
        FDS_GRAM_MMA_CASE(5); FDS_GRAM_MMA_CASE(6); FDS_GRAM_MMA_CASE(7);
    }
#undef FDS_GRAM_MMA_CASE
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------ //
//   apply_mma   out_i = A X_i for every site, A the constant (m+1) x (m+2) matrix of the boundary
//               (projection, L^-1, b^T L^-1 ...; lower triangular in its first m+1 columns plus the d
//               column).  As a GEMM: Out[sites x outputs] = X[sites x MX] A^T; a warp takes 8 sites per
//               group, the A operand fragments come straight from the TMA-staged raw rows (FD tangents
//               formed in registers), and the B fragments -- A^T, the same for every site -- are loaded
//               ONCE into registers (only the blocks the triangular structure leaves non-zero).  The
//               result goes through a shared-memory tile so that the explicit tangents and / or the next
//               perturbed ensemble u + eps * out leave as coalesced rows.  ~110 registers, two CTAs per
//               SM: one CTA's stores overlap the other's loads and tensor-core work.
// ------------------------------------------------------------------------------------------------ //
constexpr int kApplyMmaWarps = 7;       // consumer warps: 7 + 1 producer = 256 threads, two CTAs per SM at 128 registers
constexpr int kApplyMmaSites = 112;     // sites per tile: 14 groups of 8, two per consumer warp

template <int MX>
struct ApplyMmaShape {
    static constexpr int MO = MX - 1, D = MX - 1;
    static constexpr int NJ = (MO + 7) / 8;      // output tiles of 8
    static constexpr int KS = (MX + 3) / 4;      // k-steps of 4 columns of X
    static constexpr int KD = D / 4;             // the k-step that holds the d column
    // k-step kk contributes to output tile j?  (columns l <= 8j + 7, or the d column)
    static __host__ __device__ constexpr bool used(int j, int kk) { return kk <= 2 * j + 1 || kk == KD; }
};

template <int MX, int SRC>
__global__ void __launch_bounds__(32 * (kApplyMmaWarps + 1), 2) apply_mma_kernel(
    const double* __restrict__ E, const double* Tin, const double* __restrict__ d, const double* __restrict__ A,
    double inv_eps, long long n, double* Tout, double* E_out, const double* __restrict__ u, double eps, int nst,
    BatchDesc bd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (bd.gate != nullptr && bd.gate[(size_t)blockIdx.y * bd.gate_stride] != bd.gate_want) return;      // the untaken CholeskyQR variant
    using SH = ApplyMmaShape<MX>;
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        E += off * MX; Tin += off * (MX - 1);
        if (Tout != nullptr) Tout += off * (MX - 1);
        if (d != nullptr) d += off;
        if (E_out != nullptr) { E_out += off * MX; u += off; }
        A += (size_t)blockIdx.y * bd.small_stride;
    }
    constexpr int MO = SH::MO, M = MX, D = SH::D, NJ = SH::NJ, KS = SH::KS, S = kApplyMmaSites, NC = 32 * kApplyMmaWarps;
    constexpr int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    constexpr int LD = (8 * NJ + 2) | 1;                      // odd row stride of the output tile (>= MO + 1)
    TileStage ts;
    ts.full = reinterpret_cast<uint64_t*>(smem_raw);
    ts.empty = ts.full + kBndMaxStages;
    double* os = reinterpret_cast<double*>(smem_raw + 128);   // [S][LD] output rows (+ u in column 8 NJ)
    ts.raw = os + (((size_t)S * LD + 15) & ~(size_t)15);
    ts.stage_doubles = (S * W + 2 * S + 15) & ~15;
    ts.nst = nst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ts.nst; ++s) { mbar_init(&ts.full[s], 1); mbar_init(&ts.empty[s], kApplyMmaWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (n + S - 1) / S;
    if (warp == kApplyMmaWarps) { tile_producer<SRC>(ts, E, Tin, d, E_out != nullptr ? u : nullptr, MX, n, S, ntiles); return; }

    const int g = lane >> 2, t = lane & 3;
    // B fragments: b[j][kk] = A[o = 8j + g][l = 4kk + t]   (col-major 4 x 8 operand), zero outside A
    double bf[NJ][KS];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const int o = 8 * j + g, l = 4 * kk + t;
            bf[j][kk] = (SH::used(j, kk) && o < MO && l < MX) ? A[o * MX + l] : 0.0;
        }

    int st = 0;
    uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int cnt = (int)((n - s0) < S ? (n - s0) : S);
        mbar_wait(&ts.full[st], par);
        const double* raw = ts.raw + (size_t)st * ts.stage_doubles;
#pragma unroll 1
        for (int grp = warp; grp < S / 8; grp += kApplyMmaWarps) {
            const int site = 8 * grp + g;
            const double* row = raw + (size_t)site * W;
            const double e0 = SRC == SRC_ENSEMBLE ? row[0] : 0.0;
            double xa[KS];                                     // A operand: X[site g][4kk + t]
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int l = 4 * kk + t;
                double v = 0.0;
                if (l < MO) v = SRC == SRC_ENSEMBLE ? (row[l + 1] - e0) * inv_eps : row[l];
                else if (l == D) v = d != nullptr ? raw[(size_t)S * W + site] : 0.0;
                xa[kk] = v;
            }
            double acc[NJ][2];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                acc[j][0] = 0.0; acc[j][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                    if (SH::used(j, kk)) dmma_8x8x4(acc[j][0], acc[j][1], xa[kk], bf[j][kk]);
            }
            // C fragment: row = site g, columns 8j + 2t + {0, 1}
            double* orow = os + site * LD + 2 * t;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { orow[8 * j] = acc[j][0]; orow[8 * j + 1] = acc[j][1]; }
            if (E_out != nullptr && t == 0) os[site * LD + 8 * NJ] = raw[(size_t)S * W + S + site];   // u_i
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);              // the raw stage may be refilled
        if (++st == ts.nst) { st = 0; par ^= 1u; }
        named_barrier(1, NC);
        // coalesced stores: explicit tangents and / or the next perturbed ensemble
        // the rows of a tile are one contiguous block of the output arrays: flat, fully coalesced
        if (Tout != nullptr) {
            double* dst = Tout + s0 * MO;
            for (int e = tid; e < cnt * MO; e += NC) {
                const int s = e / MO, o = e - s * MO;
                dst[e] = os[s * LD + o];
            }
        }
        if (E_out != nullptr) {
            double* dst = E_out + s0 * M;
            for (int e = tid; e < cnt * M; e += NC) {
                const int s = e / M, k = e - s * M;
                const double ui = os[s * LD + 8 * NJ];
                dst[e] = k == 0 ? ui : __dadd_rn(ui, __dmul_rn(os[s * LD + k - 1], eps));
            }
        }
        named_barrier(1, NC);
    }
}

template <int MX>
static cudaError_t apply_mma_launch(int src, const double* E, const double* Tin, const double* d, const double* A,
                                    double inv_eps, long long n, double* Tout, double* E_out, const double* u,
                                    double eps, int num_sms, cudaStream_t st, LaunchStats* ls, const BatchDesc& bd) {
    using SH = ApplyMmaShape<MX>;
    constexpr int S = kApplyMmaSites, LD = (8 * SH::NJ + 2) | 1;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + 2 * S + 15) & ~(size_t)15;
    const size_t fixed = 128 + sizeof(double) * (((size_t)S * LD + 15) & ~(size_t)15);
    const int nst = (int)std::min<size_t>(kBndMaxStages, (112 * 1024 - fixed) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = fixed + sizeof(double) * nst * stage_doubles;
    const long long ntiles = (n + S - 1) / S;
    const dim3 grid((unsigned)std::min<long long>(ntiles, bd.count > 1 ? 4LL : 2LL * num_sms), (unsigned)bd.count);
    const int block = 32 * (kApplyMmaWarps + 1);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(apply_mma_kernel<MX, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_mma_kernel<MX, SRC_ENSEMBLE><<<grid, block, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst, bd);
    } else {
        e = cudaFuncSetAttribute(apply_mma_kernel<MX, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_mma_kernel<MX, SRC_TANGENT><<<grid, block, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst, bd);
    }
    if (ls) ls->launches++;
    return cudaGetLastError();
}


// The same pass for m a multiple of 8 (MX = m + 2) with everything that does not fill a tensor-core tile taken off
// the tensor cores.  A is lower triangular in its first m columns for the m homogeneous outputs; what remains is
//   * the d column: out[o] += A[o][m+1] d_i for o < m          -- 8 DFMA per lane on the C fragment,
//   * the inhomogeneous row: out[m] = sum_l A[m][l] X_i[l]      -- 8 DFMA per lane on the A fragments it already
//     holds, folded over the 4 lanes of a site by two shuffles, + the w and d terms.
// The padded tiles they occupied (one output tile with 1 live row, one k-step with 2 live columns) cost 13 of the
// 33 DMMA of a group at m = 32; DMMA runs at the vector pipe's FMA rate here, so this is a third of the pipe time.
template <int MX, int SRC>
__global__ void __launch_bounds__(32 * (kApplyMmaWarps + 1), 2) apply_mma2_kernel(
    const double* __restrict__ E, const double* Tin, const double* __restrict__ d, const double* __restrict__ A,
    double inv_eps, long long n, double* Tout, double* E_out, const double* __restrict__ u, double eps, int nst,
    BatchDesc bd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (bd.gate != nullptr && bd.gate[(size_t)blockIdx.y * bd.gate_stride] != bd.gate_want) return;      // the untaken CholeskyQR variant
    {   // problem blockIdx.y of a batch
        const size_t off = (size_t)blockIdx.y * bd.slot;
        E += off * MX; Tin += off * (MX - 1);
        if (Tout != nullptr) Tout += off * (MX - 1);
        if (d != nullptr) d += off;
        if (E_out != nullptr) { E_out += off * MX; u += off; }
        A += (size_t)blockIdx.y * bd.small_stride;
    }
    constexpr int m = MX - 2, MO = MX - 1, M = MX, D = MX - 1, NJ = m / 8, KS = m / 4;
    static_assert(m % 8 == 0, "apply_mma2: m must be a multiple of 8");
    constexpr int S = kApplyMmaSites, NC = 32 * kApplyMmaWarps;
    constexpr int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    constexpr int LD = (8 * NJ + 2) | 1;                      // odd row stride of the output tile: m outputs, w, u
    TileStage ts;
    ts.full = reinterpret_cast<uint64_t*>(smem_raw);
    ts.empty = ts.full + kBndMaxStages;
    double* os = reinterpret_cast<double*>(smem_raw + 128);   // [S][LD] output rows: [0, m) homogeneous, m: w, m + 1: u
    ts.raw = os + (((size_t)S * LD + 15) & ~(size_t)15);
    ts.stage_doubles = (S * W + 2 * S + 15) & ~15;
    ts.nst = nst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ts.nst; ++s) { mbar_init(&ts.full[s], 1); mbar_init(&ts.empty[s], kApplyMmaWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long ntiles = (n + S - 1) / S;
    // from the ensemble the primal u_i is column 0 of the staged row itself: no separate copy of u
    if (warp == kApplyMmaWarps) {
        tile_producer<SRC>(ts, E, Tin, d, (E_out != nullptr && SRC != SRC_ENSEMBLE) ? u : nullptr, MX, n, S, ntiles);
        return;
    }

    const int g = lane >> 2, t = lane & 3;
    // B fragments of the triangular part: b[j][kk] = A[o = 8j + g][l = 4kk + t], kk <= 2j + 1
    double bf[NJ][KS];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) bf[j][kk] = kk <= 2 * j + 1 ? A[(8 * j + g) * MX + 4 * kk + t] : 0.0;
    // d column on the C fragment (outputs 8j + 2t + {0, 1}) and the inhomogeneous row on the A fragments (l = 4kk + t)
    double ad[NJ][2], aw[KS];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { ad[j][0] = A[(8 * j + 2 * t) * MX + D]; ad[j][1] = A[(8 * j + 2 * t + 1) * MX + D]; }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) aw[kk] = A[m * MX + 4 * kk + t];
    const double aww = A[m * MX + m], awd = A[m * MX + D];

    int st = 0;
    uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile * S;
        const int cnt = (int)((n - s0) < S ? (n - s0) : S);
        mbar_wait(&ts.full[st], par);
        const double* raw = ts.raw + (size_t)st * ts.stage_doubles;
#pragma unroll 1
        for (int grp = warp; grp < S / 8; grp += kApplyMmaWarps) {
            const int site = 8 * grp + g;
            const double* row = raw + (size_t)site * W;
            const double e0 = SRC == SRC_ENSEMBLE ? row[0] : 0.0;
            const double dv = d != nullptr ? raw[(size_t)S * W + site] : 0.0;
            const double wv = SRC == SRC_ENSEMBLE ? (row[m + 1] - e0) * inv_eps : row[m];
            double xa[KS];                                     // A operand: X[site g][4kk + t]
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int l = 4 * kk + t;
                xa[kk] = SRC == SRC_ENSEMBLE ? (row[l + 1] - e0) * inv_eps : row[l];
            }
            double acc[NJ][2];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                acc[j][0] = ad[j][0] * dv; acc[j][1] = ad[j][1] * dv;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk)
                    if (kk <= 2 * j + 1) dmma_8x8x4(acc[j][0], acc[j][1], xa[kk], bf[j][kk]);
            }
            double wo = 0.0;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) wo = fma(aw[kk], xa[kk], wo);
            wo += __shfl_xor_sync(0xffffffffu, wo, 1);
            wo += __shfl_xor_sync(0xffffffffu, wo, 2);
            // C fragment: row = site g, columns 8j + 2t + {0, 1}
            double* orow = os + site * LD + 2 * t;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { orow[8 * j] = acc[j][0]; orow[8 * j + 1] = acc[j][1]; }
            if (t == 0) {
                os[site * LD + m] = fma(awd, dv, fma(aww, wv, wo));
                if (E_out != nullptr) os[site * LD + m + 1] = SRC == SRC_ENSEMBLE ? e0 : raw[(size_t)S * W + S + site];   // u_i
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.empty[st]);              // the raw stage may be refilled
        if (++st == ts.nst) { st = 0; par ^= 1u; }
        named_barrier(1, NC);
        // the rows of a tile are one contiguous block of the output arrays: flat, fully coalesced stores
        if (Tout != nullptr) {
            double* dst = Tout + s0 * MO;
            for (int e = tid; e < cnt * MO; e += NC) {
                const int s = e / MO, o = e - s * MO;
                dst[e] = os[s * LD + o];
            }
        }
        if (E_out != nullptr) {
            double* dst = E_out + s0 * M;
            for (int e = tid; e < cnt * M; e += NC) {
                const int s = e / M, k = e - s * M;
                const double ui = os[s * LD + m + 1];
                dst[e] = k == 0 ? ui : __dadd_rn(ui, __dmul_rn(os[s * LD + k - 1], eps));
            }
        }
        named_barrier(1, NC);
    }
}

template <int MX>
static cudaError_t apply_mma2_launch(int src, const double* E, const double* Tin, const double* d, const double* A,
                                     double inv_eps, long long n, double* Tout, double* E_out, const double* u,
                                     double eps, int num_sms, cudaStream_t st, LaunchStats* ls, const BatchDesc& bd) {
    constexpr int S = kApplyMmaSites, LD = (MX - 2 + 2) | 1;
    const int W = src == SRC_ENSEMBLE ? MX : MX - 1;
    const size_t stage_doubles = ((size_t)S * W + 2 * S + 15) & ~(size_t)15;
    const size_t fixed = 128 + sizeof(double) * (((size_t)S * LD + 15) & ~(size_t)15);
    const int nst = (int)std::min<size_t>(kBndMaxStages, (112 * 1024 - fixed) / (stage_doubles * sizeof(double)));
    if (nst < 2) return cudaErrorInvalidValue;
    const size_t smem = fixed + sizeof(double) * nst * stage_doubles;
    const long long ntiles = (n + S - 1) / S;
    const dim3 grid((unsigned)std::min<long long>(ntiles, bd.count > 1 ? 4LL : 2LL * num_sms), (unsigned)bd.count);
    const int block = 32 * (kApplyMmaWarps + 1);
    cudaError_t e;
    if (src == SRC_ENSEMBLE) {
        e = cudaFuncSetAttribute(apply_mma2_kernel<MX, SRC_ENSEMBLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_mma2_kernel<MX, SRC_ENSEMBLE><<<grid, block, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst, bd);
    } else {
        e = cudaFuncSetAttribute(apply_mma2_kernel<MX, SRC_TANGENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        apply_mma2_kernel<MX, SRC_TANGENT><<<grid, block, smem, st>>>(E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, nst, bd);
    }
    if (ls) ls->launches++;
    return cudaGetLastError();
}

static bool apply_split_on() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("FDS_APPLY_SPLIT"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

bool apply_mma_supported(int m) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("FDS_APPLY_MMA"); on = (e && e[0] == '0') ? 0 : 1; }
    if (!on) return false;
    switch (m + 2) { case 18: case 34: return true; }
    return false;
}

// A: the (m+1) x (m+2) matrix in DEVICE memory
cudaError_t launch_apply_mma(int src, const double* E, const double* Tin, const double* d, const double* A, int m,
                             double inv_eps, long long n, double* Tout, double* E_out, const double* u, double eps,
                             int num_sms, cudaStream_t st, LaunchStats* ls, const BatchDesc* batch) {
    BatchDesc bd;
    if (batch) bd = *batch;
    if (apply_split_on()) {
        switch (m + 2) {
            case 18: return apply_mma2_launch<18>(src, E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, num_sms, st, ls, bd);
            case 34: return apply_mma2_launch<34>(src, E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, num_sms, st, ls, bd);
        }
    }
    switch (m + 2) {
        case 18: return apply_mma_launch<18>(src, E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, num_sms, st, ls, bd);
        case 34: return apply_mma_launch<34>(src, E, Tin, d, A, inv_eps, n, Tout, E_out, u, eps, num_sms, st, ls, bd);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fds
