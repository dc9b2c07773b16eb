// PTX wrappers shared by the tcgen05 kernels that use the no-swizzle K-major operand layout
// (rf_tc_conv_halo.cu, rf_tc_mlp.cu): mbarriers, bulk copies, tcgen05 fences / commit / MMA / TMEM loads.
#pragma once
#include <cuda_fp16.h>

#include "rf_common.cuh"

namespace rf_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Whole-warp wait with a warp-uniform loop condition (a vote): the code after it stays provably convergent, which
// the compiler needs in order to keep an MMA issue loop on the uniform datapath.  A pipeline bug must never hang
// the GPU: every wait traps after a bounded number of polls.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if (++spins > (1u << 26)) __trap();
    }
}
// same with back-off: a few fast polls (the common short wait), then sleep between polls so that a warp that waits
// for thousands of cycles (an issuer waiting for the epilogue, a weight block in flight) does not take issue slots
// and mbarrier bandwidth from the warps that are working
__device__ __forceinline__ void mbar_wait_warp_backoff(uint32_t bar, uint32_t parity, unsigned ns = 64) {
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if (++spins > 8) __nanosleep(ns);
        if (spins > (1u << 24)) __trap();
    }
}
// same, polling gently: for warps that wait for a whole phase of MMAs
__device__ __forceinline__ void mbar_wait_warp_sleepy(uint32_t bar, uint32_t parity, unsigned ns = 100) {
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        __nanosleep(ns);
        if (++spins > (1u << 24)) __trap();
    }
}
// single producer thread: it waits for a long time, polling at full speed would take shared-memory cycles away from
// the tensor core's operand reads
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns = 200) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        __nanosleep(ns);
        if (spins > (1u << 23)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// tcgen05.mma (M128, kind::f16, cta_group::1) with the shared-memory descriptors passed as (low, high) words.
//   low  = start address >> 4 | (LBO >> 4) << 16     (the only word that changes between MMAs)
//   high = SBO >> 4 | 1 << 14 (descriptor version)   (no-swizzle K-major: core matrix = 8 rows x 16 B, LBO = byte
//          distance between the two K chunks of a K = 16 step, SBO = distance between consecutive 8-row groups)
// `issue` (1 on the elected lane) predicates the instruction INSIDE the asm block: with a C++ `if (leader)` around
// it the compiler sinks the descriptor arithmetic into the divergent region.
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc, uint32_t issue) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(issue)
        : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
// Issue a 32-lane x 16-column TMEM load without waiting (pair with tc_ld_wait): several loads in flight per thread.
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Ties 16 registers that an earlier tc_ld16_issue filled to this point of the program: arithmetic on them cannot be
// scheduled above it (place it right after tc_ld_wait when other work sits between the issue and the wait).
__device__ __forceinline__ void tc_ld_fence16(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                 "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}
// instruction descriptor: D f32, A/B f16, both K-major, N >> 3 at bit 17, M (128) >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
// x = hi + lo + r with hi, lo fp16: |r| <= 2^-24 |x| for |x| in fp16's normal range (the operands here are
// GroupNorm-ed activations, TSDF patches and weights, all O(1)); hi is saturated so that even |x| up to 1.3e5
// splits without producing inf.
__device__ __forceinline__ void split_f16(float x, uint32_t& hi, uint32_t& lo) {
    const __half h = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// Two values at once with the packed converts (F2FP.PACK_AB / HADD2.F32): ~5 instructions per value instead of ~8.
// Same result as split_f16 on each element.
// The saturation is part of the conversion (cvt.rn.satfinite.f16x2.f32 = F2FP.SATFINITE...PACK_AB): for finite inputs the
// same bits as clamping to +-65504 first, without the four FMNMX per pair that made up a quarter of the conversion
// loops' instructions (the lo part saturates the same way, so |x| up to 1.3e5 still splits without inf).
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - hf.y), "f"(x0 - hf.x));
}

}  // namespace rf_tc
