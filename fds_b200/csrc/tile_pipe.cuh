// TMA-fed tile pipeline shared by the boundary kernels (boundary_fast.cu, boundary_mma.cu).
#pragma once
#include "kernels.h"
#include "common.cuh"

namespace fds {

// ------------------------- TMA-fed tile pipeline ----------------------------- //
// A tile is S consecutive sites: the rows [s0, s0+S) of the source array (ensemble rows of
// M = m+2 doubles, or tangent rows of MO = m+1 doubles) are ONE contiguous block of global
// memory, fetched with a single cp.async.bulk, plus S doubles of d.  A dedicated producer
// warp keeps `NST` stages in flight; consumer warps release a stage as soon as they have
// pulled its contents into registers / the work buffer.
constexpr int kBndMaxStages = 4;

struct TileStage {
    uint64_t* full;
    uint64_t* empty;
    double* raw;          // stage 0: [S][W] rows, then [S] d values, then [S] u values
    int stage_doubles;    // doubles per stage
    int nst;              // stages in the ring
};

template <int SRC>
FDS_D void tile_producer(const TileStage& ts, const double* __restrict__ E, const double* __restrict__ T,
                         const double* __restrict__ d, const double* __restrict__ u, int MX, long long n, int S,
                         long long ntiles) {
    if ((threadIdx.x & 31) != 0) return;
    const int W = SRC == SRC_ENSEMBLE ? MX : MX - 1;
    const double* src = SRC == SRC_ENSEMBLE ? E : T;
    int st = 0;
    uint32_t par = 0;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        if (it >= ts.nst) mbar_wait(&ts.empty[st], par ^ 1u);
        const long long s0 = tile * S;
        const long long cnt = (n - s0) < S ? (n - s0) : S;          // sites in this tile (even, or the last tile)
        // bulk copies need 16-byte sizes: an odd trailing double is fetched by rounding the byte
        // count up (the arrays are padded by the allocator; see ctx.cu)
        const uint32_t rb = (uint32_t)(((size_t)cnt * W * 8 + 15) & ~(size_t)15);
        const uint32_t db = d != nullptr ? (uint32_t)(((size_t)cnt * 8 + 15) & ~(size_t)15) : 0u;
        double* dst = ts.raw + (size_t)st * ts.stage_doubles;
        const uint32_t ub = u != nullptr ? (uint32_t)(((size_t)cnt * 8 + 15) & ~(size_t)15) : 0u;
        mbar_expect_tx(&ts.full[st], rb + db + ub);
        bulk_g2s(dst, src + s0 * W, rb, &ts.full[st]);
        if (d != nullptr) bulk_g2s(dst + (size_t)S * W, d + s0, db, &ts.full[st]);
        if (u != nullptr) bulk_g2s(dst + (size_t)S * W + S, u + s0, ub, &ts.full[st]);
        if (++st == ts.nst) { st = 0; par ^= 1u; }
    }
}

}  // namespace fds
