// LssTangent.solve on the device (K7 of SURVEY.md 2.3).
//
// Reference fds/lsstan.py:36-50: minimise sum |alpha_i|^2 subject to
// alpha_{i+1} = R_i alpha_i + b_i  (i = 0..K-1), i.e. B alpha = -b... solved there
// through the Schur complement S = B B^T with scipy SuperLU.  Here: S is block
// tridiagonal SPD with  S_ii = R_i R_i^T + I,  S_{i,i-1} = -R_i,  so a block
// Cholesky (block Thomas) sweep in ONE CTA does it:
//     E_i = S_{i,i-1} C_{i-1}^{-T},  C_i C_i^T = S_ii - E_i E_i^T,
//     y_i = C_i^{-1} (b_i - E_i y_{i-1}),   lambda_i = C_i^{-T} (y_i - E_{i+1}^T lambda_{i+1}),
//     alpha_i = lambda_{i-1} - R_i^T lambda_i.
// K is sequential by nature; the m x m block operations are spread over the CTA.
#include "kernels.h"
#include "common.cuh"

namespace fds {

size_t lss_smem_bytes(int m) { return sizeof(double) * (4 * (size_t)m * m + 4 * m) + 16; }

__global__ void __launch_bounds__(512) lss_solve_kernel(const double* __restrict__ R, const double* __restrict__ b,
                                                        long long K, int m, double* __restrict__ alpha,
                                                        double* __restrict__ work, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                 // m*m   current diagonal block -> its Cholesky factor C_i
    double* Cp = S + m * m;         // m*m   previous factor C_{i-1}
    double* Eb = Cp + m * m;        // m*m   E_i
    double* Rb = Eb + m * m;        // m*m   R_i
    double* y = Rb + m * m;         // m     y_i
    double* yp = y + m;             // m     y_{i-1} / lambda_{i+1}
    double* rhs = yp + m;           // m
    double* tmp = rhs + m;          // m
    __shared__ int fail;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t mm = (size_t)m * m;
    {   // problem blockIdx.x of a batch
        const size_t p = blockIdx.x;
        R += p * K * mm; b += p * K * m; alpha += p * K * m;
        work += p * ((size_t)2 * K * mm + (size_t)2 * K * m);
        status += p;
    }
    double* Cst = work;                       // [K][m][m]
    double* Est = work + (size_t)K * mm;      // [K][m][m]
    double* yst = Est + (size_t)K * mm;       // [K][m]
    double* lst = yst + (size_t)K * m;        // [K][m]
    if (tid == 0) fail = 0;
    __syncthreads();

    // ------------------------------ forward sweep ------------------------------ //
    for (long long i = 0; i < K; ++i) {
        for (int e = tid; e < (int)mm; e += nthr) Rb[e] = R[i * mm + e];
        __syncthreads();
        // S = R_i R_i^T + I  (lower triangle is enough)
        for (int e = tid; e < (int)mm; e += nthr) {
            const int r = e / m, c = e % m;
            double s = (r == c) ? 1.0 : 0.0;
            if (c <= r) for (int k = 0; k < m; ++k) s = fma(Rb[r * m + k], Rb[c * m + k], s);
            S[e] = s;
        }
        for (int r = tid; r < m; r += nthr) rhs[r] = b[i * m + r];
        __syncthreads();
        if (i > 0) {
            // E_i C_{i-1}^T = -R_i : row r of E by forward substitution against C_{i-1}
            for (int r = tid; r < m; r += nthr) {
                for (int c = 0; c < m; ++c) {
                    double s = -Rb[r * m + c];
                    for (int k = 0; k < c; ++k) s -= Eb[r * m + k] * Cp[c * m + k];
                    Eb[r * m + c] = s / Cp[c * m + c];
                }
            }
            __syncthreads();
            for (int e = tid; e < (int)mm; e += nthr) {
                const int r = e / m, c = e % m;
                if (c <= r) {
                    double s = S[e];
                    for (int k = 0; k < m; ++k) s -= Eb[r * m + k] * Eb[c * m + k];
                    S[e] = s;
                }
            }
            for (int r = tid; r < m; r += nthr) {
                double s = rhs[r];
                for (int k = 0; k < m; ++k) s -= Eb[r * m + k] * yp[k];
                tmp[r] = s;
            }
            __syncthreads();
            for (int r = tid; r < m; r += nthr) rhs[r] = tmp[r];
            for (int e = tid; e < (int)mm; e += nthr) Est[i * mm + e] = Eb[e];
            __syncthreads();
        }
        // Cholesky S = C C^T in place (lower)
        for (int j = 0; j < m; ++j) {
            __syncthreads();
            const double piv = S[j * m + j];
            if (!(piv > 0.0) && tid == 0 && fail == 0) fail = (int)(i * m + j + 1);
            const double ljj = sqrt(piv > 0.0 ? piv : 1.0);
            __syncthreads();
            if (tid == 0) S[j * m + j] = ljj;
            for (int r = j + 1 + tid; r < m; r += nthr) S[r * m + j] /= ljj;
            __syncthreads();
            const int rem = m - j - 1;
            for (int e = tid; e < rem * rem; e += nthr) {
                const int r = j + 1 + e / rem, c = j + 1 + e % rem;
                if (c <= r) S[r * m + c] -= S[r * m + j] * S[c * m + j];
            }
        }
        __syncthreads();
        // y_i = C_i^{-1} rhs (sequential forward substitution, one thread; m is small)
        if (tid == 0) {
            for (int r = 0; r < m; ++r) {
                double s = rhs[r];
                for (int k = 0; k < r; ++k) s -= S[r * m + k] * y[k];
                y[r] = s / S[r * m + r];
            }
        }
        __syncthreads();
        for (int e = tid; e < (int)mm; e += nthr) { Cst[i * mm + e] = S[e]; Cp[e] = S[e]; }
        for (int r = tid; r < m; r += nthr) { yst[i * m + r] = y[r]; yp[r] = y[r]; }
        __syncthreads();
    }
    // ------------------------------ backward sweep ----------------------------- //
    for (long long i = K - 1; i >= 0; --i) {
        for (int e = tid; e < (int)mm; e += nthr) S[e] = Cst[i * mm + e];
        for (int r = tid; r < m; r += nthr) rhs[r] = yst[i * m + r];
        if (i + 1 < K) for (int e = tid; e < (int)mm; e += nthr) Eb[e] = Est[(i + 1) * mm + e];
        __syncthreads();
        if (i + 1 < K) {
            for (int r = tid; r < m; r += nthr) {          // rhs -= E_{i+1}^T lambda_{i+1}
                double s = rhs[r];
                for (int k = 0; k < m; ++k) s -= Eb[k * m + r] * yp[k];
                tmp[r] = s;
            }
            __syncthreads();
            for (int r = tid; r < m; r += nthr) rhs[r] = tmp[r];
            __syncthreads();
        }
        if (tid == 0) {                                    // lambda_i = C_i^{-T} rhs
            for (int r = m - 1; r >= 0; --r) {
                double s = rhs[r];
                for (int k = r + 1; k < m; ++k) s -= S[k * m + r] * y[k];
                y[r] = s / S[r * m + r];
            }
        }
        __syncthreads();
        for (int r = tid; r < m; r += nthr) { lst[i * m + r] = y[r]; yp[r] = y[r]; }
        __syncthreads();
    }
    // alpha_i = lambda_{i-1} - R_i^T lambda_i
    for (long long e = tid; e < K * m; e += nthr) {
        const long long i = e / m;
        const int r = (int)(e % m);
        double s = i > 0 ? lst[(i - 1) * m + r] : 0.0;
        for (int k = 0; k < m; ++k) s -= R[i * mm + (size_t)k * m + r] * lst[i * m + k];
        alpha[e] = s;
    }
    __syncthreads();
    if (tid == 0) *status = fail;
}

cudaError_t launch_lss_solve(const double* R, const double* b, long long K, int m, double* alpha, double* work,
                             int* status, cudaStream_t st, LaunchStats* ls, int batch) {
    const size_t smem = lss_smem_bytes(m);
    cudaError_t e = cudaFuncSetAttribute(lss_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lss_solve_kernel<<<batch > 0 ? batch : 1, 512, smem, st>>>(R, b, K, m, alpha, work, status);
    if (ls) ls->launches++;
    return cudaGetLastError();
}

}  // namespace fds
