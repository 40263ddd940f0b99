// Common device helpers for the EXAUQ B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EXQ_TILE 128      // all dense matrices are padded to a multiple of this
#define EXQ_MAX_D 64      // max input dimension supported by the kernels

// kernel ids (order mirrors exauq/core/emulators.py:122-126)
#define EXQ_KERNEL_MATERN52 0
#define EXQ_KERNEL_SQEXP 1
#define EXQ_KERNEL_PRODMAT52 2

namespace exq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- Ampere-style 16-byte async copy (LDGSTS) ------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef EXQ_CP_ASYNC_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- TMA 1-D bulk copy (cp.async.bulk -> UBLKCP) with an mbarrier ------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}
// TMA tiled load of one 3-D box (coordinates innermost first) into shared memory
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col) -> SASS DMMA.8x8x4 ----
// lane = 4*g + t:  a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- covariance functions (mogp-emulator Kernel.py; call site exauq/core/emulators.py:122-126)
// r2 = sum_k exp(corr_raw_k) (x_k - x'_k)^2
__device__ __forceinline__ double corr_from_r2(int kernel_id, double r2) {
    if (kernel_id == EXQ_KERNEL_MATERN52) {
        double a = sqrt(5.0 * r2);
        return (1.0 + a + (5.0 / 3.0) * r2) * exp(-a);
    }
    return exp(-0.5 * r2);
}
// d corr / d r2
__device__ __forceinline__ double dcorr_dr2(int kernel_id, double r2) {
    if (kernel_id == EXQ_KERNEL_MATERN52) {
        double a = sqrt(5.0 * r2);
        return (-5.0 / 6.0) * (1.0 + a) * exp(-a);
    }
    return -0.5 * exp(-0.5 * r2);
}
// ---- branch-free FP64 exp / sqrt for the assembly and gradient kernels ------------------------
// The library exp() / sqrt() carry slow-path branches (range checks, denormal fix-ups); with 32
// entries per thread those branches keep the compiler from interleaving the 32 independent
// evaluation chains, and the assembly kernel ran its FP64 pipe at 44 % (ncu).  The arguments
// here are known: exp of a non-positive number, sqrt of a non-negative one.
// exp(x) for x <= 0, error < 1 ulp; results below 2^-995 are returned as ~1e-300 (never compared: a
// correlation that small multiplies a covariance into the denormal range either way)
__device__ __forceinline__ double exp_nonpos(double x) {
    x = fmax(x, -690.0);
    const double SHIFT = 6755399441055744.0;                 // 1.5 * 2^52: rounds to an integer in the low word
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const int k = __double2loint(t);
    const double kd = t - SHIFT;
    double r = fma(kd, -6.93147180369123816490e-01, x);      // ln2 split: high part exact for |k| < 2^20
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.3888888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));   // p * 2^k, k >= -996
}
// sqrt(x) for 0 <= x, within 1 ulp; x is clamped to [1e-30, 1e30] (below: the Matern factor is 1.0 to
// 1e-30 either way; above: exp(-sqrt(x)) is 0 either way)
__device__ __forceinline__ double sqrt_nonneg(double x) {
    const double xs = fmin(fmax(x, 1e-30), 1e30);
    double y = (double)rsqrtf((float)xs);                    // 23 bits
    const double h = 0.5 * xs;
    y = y * fma(-h * y, y, 1.5);                             // 46 bits
    y = y * fma(-h * y, y, 1.5);                             // > 53 bits
    double s = xs * y;
    return fma(fma(-s, s, xs), 0.5 * y, s);                  // final correction of s = sqrt(xs)
}
// corr(r2) with the kernel fixed at compile time
template <int KID>
__device__ __forceinline__ double corr_from_r2_t(double r2) {
    if (KID == EXQ_KERNEL_MATERN52) {
        const double a = sqrt_nonneg(5.0 * r2);
        return (1.0 + a + (5.0 / 3.0) * r2) * exp_nonpos(-a);
    }
    return exp_nonpos(-0.5 * r2);
}
template <int KID>
__device__ __forceinline__ double mat52_1d_t(double r2) {
    const double a = sqrt_nonneg(5.0 * r2);
    return (1.0 + a + (5.0 / 3.0) * r2) * exp_nonpos(-a);
}

// 1-D Matern-5/2 factor for the product kernel
__device__ __forceinline__ double mat52_1d(double r2) {
    double a = sqrt(5.0 * r2);
    return (1.0 + a + (5.0 / 3.0) * r2) * exp(-a);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace exq
