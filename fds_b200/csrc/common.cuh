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

#if defined(__CUDACC__)
// ------------- PTX wrappers: mbarrier + 1-D TMA bulk copy (cp.async.bulk) ------------- //
FDS_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FDS_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
FDS_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
FDS_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
FDS_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
FDS_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// the same on a 32-bit shared-window address computed once (no generic->shared conversion in the loop)
FDS_D void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (ok == 0);
}
// non-blocking probe of a phase (never suspends the thread)
FDS_D bool mbar_test_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
FDS_D void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 8-byte load from a shared-window address (ordered after the mbarrier waits: all are volatile)
FDS_D double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
FDS_D void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
template <int OFF>
FDS_D void sts_f64_off(uint32_t a, double v) { asm volatile("st.shared.f64 [%0+%2], %1;" ::"r"(a), "d"(v), "n"(OFF)); }
FDS_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

FDS_D void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

}  // namespace fds
