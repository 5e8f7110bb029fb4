// FP64 issue-rate microbenchmark (development aid): how many DADD / DMUL / DFMA warp-instructions per clock
// does one SM sub-partition sustain with 1, 2 or 4 resident warps?   nvcc -O3 -arch=sm_100a -o dpbench dpbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>   // 0: DADD, 1: DMUL, 2: DFMA, 3: the RK4 mix (2 DADD : 1 DMUL)
__global__ void dp_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + j + threadIdx.x * 1e-3;
    const double c = seed * 0.999, d = seed * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) a[j] = __dadd_rn(a[j], d);
                else if (KIND == 1) a[j] = __dmul_rn(a[j], c);
                else if (KIND == 2) a[j] = __fma_rn(a[j], c, d);
                else { a[j] = __dadd_rn(a[j], d); a[j] = __dmul_rn(a[j], c); a[j] = __dadd_rn(a[j], -d); }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int threads, double* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dp_kernel<KIND><<<148, threads>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    dp_kernel<KIND><<<148, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_thread = (double)iters * 64 * (KIND == 3 ? 3 : 1);
    const double warps_per_sched = threads / 32 / 4.0;
    const double winstr_per_sched = per_thread * warps_per_sched;
    printf("%-5s %3d thr/CTA (%g warps/scheduler): %.3f ms  -> %.2f clk per warp-instruction per scheduler at %d MHz nominal\n",
           name, threads, warps_per_sched, ms, ms * 1e-3 * clk * 1e3 / winstr_per_sched, clk / 1000);
}

int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    for (int t : {128, 256, 512}) {
        run<0>("DADD", t, out); run<1>("DMUL", t, out); run<2>("DFMA", t, out); run<3>("MIX", t, out);
    }
    return 0;
}
