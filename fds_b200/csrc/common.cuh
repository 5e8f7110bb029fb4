// Common helpers for the fds_b200 CUDA sources (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define FDS_HD __host__ __device__ __forceinline__
#define FDS_D __device__ __forceinline__
#else
#define FDS_HD inline
#define FDS_D inline
#endif

namespace fds {

// --------------------------------------------------------------------------- //
// Arithmetic policies.  Exact: every operation individually rounded, in the order
// the numpy oracle (oracle/plugins.py) performs it -> bit-identical state.
// Fast: FMA contraction where the algebra allows.
// --------------------------------------------------------------------------- //
struct ExactMath {
    static constexpr bool kExact = true;
#if defined(__CUDA_ARCH__)
    static FDS_HD double add(double a, double b) { return __dadd_rn(a, b); }
    static FDS_HD double sub(double a, double b) { return __dsub_rn(a, b); }
    static FDS_HD double mul(double a, double b) { return __dmul_rn(a, b); }
    static FDS_HD double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
    // host build (emulation harness) is compiled with -ffp-contract=off
    static FDS_HD double add(double a, double b) { return a + b; }
    static FDS_HD double sub(double a, double b) { return a - b; }
    static FDS_HD double mul(double a, double b) { return a * b; }
    static FDS_HD double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif
    // a*b + c with BOTH roundings (numpy: tmp = a*b; tmp + c)
    static FDS_HD double muladd(double a, double b, double c) { return add(mul(a, b), c); }
    // 2*k + a : 2*k is exact, so one fused op gives the numpy result bit for bit
    static FDS_HD double twice_plus(double k, double a) { return fma_(2.0, k, a); }
};

struct FastMath {
    static constexpr bool kExact = false;
    static FDS_HD double add(double a, double b) { return a + b; }
    static FDS_HD double sub(double a, double b) { return a - b; }
    static FDS_HD double mul(double a, double b) { return a * b; }
#if defined(__CUDA_ARCH__)
    static FDS_HD double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
    static FDS_HD double fma_(double a, double b, double c) { return __builtin_fma(a, b, c); }
#endif
    static FDS_HD double muladd(double a, double b, double c) { return fma_(a, b, c); }
    static FDS_HD double twice_plus(double k, double a) { return fma_(2.0, k, a); }
};

}  // namespace fds
