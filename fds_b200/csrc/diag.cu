// Diagnostics that live in the library so that bench.py reports measured roofs, not estimates.
//
// fp64_issue_rate: the exact-mode Lorenz-96 step (fds/segment.py:56-64 -> l96_march2.cuh here) is bound by
// the FP64 pipe, not by HBM (two time steps per pass halve the traffic; profiles/).  Its roof is therefore the
// rate at which the SMs issue FP64 warp-instructions at the clock the board sustains under that load.  This
// kernel measures exactly that with the step kernel's own shape: one 256-thread CTA per SM (2 warps per
// scheduler), 8 independent dependency chains per thread, a 2 DADD : 1 DMUL mix, no memory traffic.
#include "kernels.h"

namespace fds {

__global__ void __launch_bounds__(256, 1) fp64_rate_kernel(double* out, int iters, double seed, unsigned long long* cycles) {
    const long long clk0 = clock64();
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + j + threadIdx.x * 1e-3;
    const double c = seed * 0.999, d = seed * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = __dadd_rn(a[j], d);
                a[j] = __dmul_rn(a[j], c);
                a[j] = __dadd_rn(a[j], -d);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (cycles != nullptr && threadIdx.x == 0) {
        atomicAdd(cycles, (unsigned long long)(clock64() - clk0));
        atomicAdd(cycles + 1, 1ULL);
    }
}

cudaError_t fp64_issue_rate(int num_sms, cudaStream_t st, double* winst_per_s, double* seconds, double* winst_per_clk_sm) {
    double* out = nullptr;
    cudaError_t e = cudaMalloc(&out, (size_t)num_sms * 256 * sizeof(double) + 16);
    if (e != cudaSuccess) return e;
    unsigned long long* cyc = reinterpret_cast<unsigned long long*>(out + (size_t)num_sms * 256);
    cudaMemsetAsync(cyc, 0, 16, st);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // long enough (~0.25 s) for the board to settle at the clock it sustains at its power limit
    const int iters = 100000;
    fp64_rate_kernel<<<num_sms, 256, 0, st>>>(out, 2000, 1.0, nullptr);
    cudaEventRecord(e0, st);
    fp64_rate_kernel<<<num_sms, 256, 0, st>>>(out, iters, 1.0, cyc);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long hc[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    const double winst = (double)iters * 8 * 8 * 3 * (256 / 32) * num_sms;
    if (seconds) *seconds = ms * 1e-3;
    if (winst_per_s) *winst_per_s = winst / (ms * 1e-3);
    // per SM and SM clock, from the kernel's own cycle counters: independent of the clock the board ran at
    if (winst_per_clk_sm) *winst_per_clk_sm = hc[1] ? (winst / num_sms) / ((double)hc[0] / (double)hc[1]) : 0.0;
    return cudaGetLastError();
}

}  // namespace fds
