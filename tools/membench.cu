// Memory-system microbenchmark for the strip-march access pattern (development aid).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }

// plain copy
__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += st) out[i] = in[i];
}

constexpr int NST = 6;
// strip pattern: CTA owns R strips; block b of strip s = 8 sites x M doubles.
// LAYOUT 0: natural  addr(s,b) = (s*nblk + b) * 8*M
// LAYOUT 1: interleaved addr(s,b) = (b*nstrips + s) * 8*M
template <int LAYOUT, bool READ, bool WRITE>
__global__ void __launch_bounds__(512, 1) strip_kernel(const double* __restrict__ X, double* __restrict__ Y, int M, int R,
                                                       int nblk, long long nstrips, int chunk_sites) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + 8;
    double* sdata = reinterpret_cast<double*>(smem_raw + 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = R * M, ncw = (ncons + 31) >> 5;
    const int CS = chunk_sites;                   // sites per stage
    const int stage_stride = R * CS * M;
    const int nit = nblk * 8 / CS;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto base = [&](long long s, long long it) -> long long {
        return LAYOUT == 0 ? (s * nit + it) * (long long)CS * M : (it * nstrips + s) * (long long)CS * M;
    };
    if (warp == ncw) {
        if (!READ) return;
        int st = 0; uint32_t par = 0;
        const uint32_t bytes = CS * M * 8;
        for (int it = 0; it < nit; ++it) {
            if (it >= NST) mbar_wait(&empty[st], par ^ 1u);
            if (lane == 0) mbar_expect_tx(&full[st], bytes * R);
            __syncwarp();
            for (int r = lane; r < R; r += 32) {
                long long s = (long long)blockIdx.x * R + r;
                bulk_g2s(sdata + (size_t)st * stage_stride + r * CS * M, X + base(s, it), bytes, &full[st]);
            }
            if (++st == NST) { st = 0; par ^= 1u; }
        }
        return;
    }
    if (tid >= ncons) {      // idle lanes of the last consumer warp still arrive
        if (READ) { int st = 0; uint32_t par = 0;
            for (int it = 0; it < nit; ++it) { mbar_wait(&full[st], par); __syncwarp(); if (lane == 0) mbar_arrive(&empty[st]); if (++st == NST) { st = 0; par ^= 1u; } } }
        return;
    }
    const int r = tid / M, k = tid - r * M;
    const long long s = (long long)blockIdx.x * R + r;
    int st = 0; uint32_t par = 0;
    double acc = 0.0;
    for (int it = 0; it < nit; ++it) {
        const double* cur = sdata + (size_t)st * stage_stride + r * CS * M + k;
        if (READ) mbar_wait(&full[st], par);
        double* yo = Y + base(s, it) + k;
        for (int p = 0; p < CS; ++p) {
            double v = READ ? cur[p * M] : (double)p;
            acc += v;
            if (WRITE) yo[p * M] = v + 1.0;
        }
        if (READ) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[st]); }
        if (++st == NST) { st = 0; par ^= 1u; }
    }
    if (acc == 1.2345) Y[0] = acc;
}

template <int LAYOUT, bool READ, bool WRITE>
float run_strip(const double* X, double* Y, int M, int R, int nblk, long long nstrips, int cs, int reps) {
    size_t smem = 128 + (size_t)NST * R * cs * M * 8;
    CK(cudaFuncSetAttribute(strip_kernel<LAYOUT, READ, WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)(nstrips / R);
    int block = ((R * M + 31) / 32) * 32 + 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) strip_kernel<LAYOUT, READ, WRITE><<<grid, block, smem>>>(X, Y, M, R, nblk, nstrips, cs);
    CK(cudaGetLastError());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) strip_kernel<LAYOUT, READ, WRITE><<<grid, block, smem>>>(X, Y, M, R, nblk, nstrips, cs);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

// strip copy with the OUTPUT staged in shared memory and written by TMA bulk stores (one contiguous
// CS*M*8-byte chunk per strip and block) instead of per-thread 8-byte stores
constexpr int NSI = 4, NSO = 2;
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory"); }
__global__ void __launch_bounds__(512, 1) strip_tma_store_kernel(const double* __restrict__ X, double* __restrict__ Y, int M, int R,
                                                                 int nblk, long long nstrips, int CS) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + 8;
    uint64_t* ofull = full + 16;
    uint64_t* oempty = full + 24;
    double* sdata = reinterpret_cast<double*>(smem_raw + 256);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = R * M, ncw = (ncons + 31) >> 5;
    const int stage_stride = R * CS * M;
    double* odata = sdata + (size_t)NSI * stage_stride;
    const int nit = nblk * 8 / CS;
    if (tid == 0) {
        for (int s = 0; s < NSI; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
        for (int s = 0; s < NSO; ++s) { mbar_init(&ofull[s], ncw); mbar_init(&oempty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto base = [&](long long s, long long it) -> long long { return (s * nit + it) * (long long)CS * M; };
    const uint32_t bytes = CS * M * 8;
    if (warp == ncw) {          // IO warp: loads ahead, stores behind
        int st = 0; uint32_t par = 0; int so = 0; uint32_t opar = 0;
        for (int it = 0; it < nit + 1; ++it) {
            if (it < nit) {
                if (it >= NSI) mbar_wait(&empty[st], par ^ 1u);
                if (lane == 0) mbar_expect_tx(&full[st], bytes * R);
                __syncwarp();
                for (int r = lane; r < R; r += 32) {
                    long long s = (long long)blockIdx.x * R + r;
                    bulk_g2s(sdata + (size_t)st * stage_stride + r * CS * M, X + base(s, it), bytes, &full[st]);
                }
                if (++st == NSI) { st = 0; par ^= 1u; }
            }
            if (it >= 1) {
                const int j = it - 1;                 // store block j
                mbar_wait(&ofull[so], opar);
                for (int r = lane; r < R; r += 32) {
                    long long s = (long long)blockIdx.x * R + r;
                    bulk_s2g(Y + base(s, j), odata + (size_t)so * stage_stride + r * CS * M, bytes);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // the stage written NSO-1 blocks ago has been read out: hand it back
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSO - 1) : "memory");
                __syncwarp();
                if (j >= NSO - 1 && lane == 0) mbar_arrive(&oempty[(so + 1) % NSO]);
                if (++so == NSO) { so = 0; opar ^= 1u; }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }
    const int tt = tid < ncons ? tid : ncons - 1;
    const int r = tt / M, k = tt - r * M;
    int st = 0; uint32_t par = 0; int so = 0; uint32_t opar = 0;
    double acc = 0.0;
    for (int it = 0; it < nit; ++it) {
        const double* cur = sdata + (size_t)st * stage_stride + r * CS * M + k;
        double* oc = odata + (size_t)so * stage_stride + r * CS * M + k;
        mbar_wait(&full[st], par);
        if (it >= NSO) mbar_wait(&oempty[so], opar ^ 1u);
        for (int p = 0; p < CS; ++p) {
            double v = cur[p * M];
            acc += v;
            oc[p * M] = v + 1.0;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(&empty[st]); mbar_arrive(&ofull[so]); }
        if (++st == NSI) { st = 0; par ^= 1u; }
        if (++so == NSO) { so = 0; opar ^= 1u; }
    }
    if (acc == 1.2345) Y[0] = acc;
}

float run_strip_tma(const double* X, double* Y, int M, int R, int nblk, long long nstrips, int cs, int reps) {
    size_t smem = 256 + (size_t)(NSI + NSO) * R * cs * M * 8;
    CK(cudaFuncSetAttribute(strip_tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (int)(nstrips / R);
    int block = ((R * M + 31) / 32) * 32 + 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) strip_tma_store_kernel<<<grid, block, smem>>>(X, Y, M, R, nblk, nstrips, cs);
    CK(cudaGetLastError());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) strip_tma_store_kernel<<<grid, block, smem>>>(X, Y, M, R, nblk, nstrips, cs);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 34;
    const int R = argc > 2 ? atoi(argv[2]) : 14;
    const int nblk = 1016;
    const long long nstrips = 148LL * R;
    const size_t n = (size_t)nstrips * nblk * 8 * M;          // doubles
    double *X, *Y;
    CK(cudaMalloc(&X, n * 8)); CK(cudaMalloc(&Y, n * 8));
    CK(cudaMemset(X, 0, n * 8)); CK(cudaMemset(Y, 0, n * 8));
    const double gb = n * 8 / 1e9;
    printf("M=%d R=%d strips=%lld array=%.2f GB\n", M, R, nstrips, gb);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int g : {148 * 8, 148 * 16, 148 * 32}) {
            copy_kernel<<<g, 512>>>((const double2*)X, (double2*)Y, n / 2);
            cudaEventRecord(e0);
            for (int i = 0; i < 5; ++i) copy_kernel<<<g, 512>>>((const double2*)X, (double2*)Y, n / 2);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("copy grid=%d: %.3f ms  %.0f GB/s (r+w)\n", g, ms, 2 * gb / ms * 1e3 / 1e3 * 1e0 * 1.0e0);
        }
    }
    { float t = run_strip_tma(X, Y, M, R, nblk, nstrips, 8, 5); printf("cs= 8 natural tma-store copy: %.3f ms %.2f TB/s (r+w)\n", t, 2 * gb / t);
      double h[4]; CK(cudaMemcpy(h, Y + 5, 32, cudaMemcpyDeviceToHost)); printf("  check %g %g\n", h[0], h[1]); }
    for (int cs : {8}) {
        float t;
        t = run_strip<0, true, false>(X, Y, M, R, nblk, nstrips, cs, 5);  printf("cs=%2d natural     read : %.3f ms %.0f GB/s\n", cs, t, gb / t);
        t = run_strip<1, true, false>(X, Y, M, R, nblk, nstrips, cs, 5);  printf("cs=%2d interleaved read : %.3f ms %.0f GB/s\n", cs, t, gb / t);
        t = run_strip<0, false, true>(X, Y, M, R, nblk, nstrips, cs, 5);  printf("cs=%2d natural     write: %.3f ms %.0f GB/s\n", cs, t, gb / t);
        t = run_strip<1, false, true>(X, Y, M, R, nblk, nstrips, cs, 5);  printf("cs=%2d interleaved write: %.3f ms %.0f GB/s\n", cs, t, gb / t);
        t = run_strip<0, true, true>(X, Y, M, R, nblk, nstrips, cs, 5);   printf("cs=%2d natural     copy : %.3f ms %.0f GB/s (r+w)\n", cs, t, 2 * gb / t);
        t = run_strip<1, true, true>(X, Y, M, R, nblk, nstrips, cs, 5);   printf("cs=%2d interleaved copy : %.3f ms %.0f GB/s (r+w)\n", cs, t, 2 * gb / t);
    }
    return 0;
}
