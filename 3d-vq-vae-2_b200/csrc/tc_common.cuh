// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (mbarrier, bulk copy, UMMA descriptors,
// TMEM loads).  Device code only; included inside namespace vq3d by the .cu files that need it.
#pragma once
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}

// sleep_ns > 0: back off between polls (a polling warp still takes issue slots from the working warps of its
// sub-partition: in the quantizer 40 % of all issued instructions were polls before the producer/consumer roles slept)
#ifndef VQ3D_MBAR_SLEEP_NS
#define VQ3D_MBAR_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbarrier_wait(uint64_t *bar, uint32_t parity, uint32_t sleep_ns = VQ3D_MBAR_SLEEP_NS) {
    const uint32_t addr = s_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        uint32_t ok;
        // the suspend-time hint lets the hardware park the warp until the phase completes instead of returning early
        // and spinning (spinning warps take issue slots from the working warps of the same sub-partition)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
        if (ok) return;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
    __trap();   // a lost arrival becomes a launch error, never a hung GPU
}

__device__ __forceinline__ bool mbarrier_test(uint64_t *bar, uint32_t parity) {     // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbarrier_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbarrier_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy (UBLKCP); completion is signalled as tx bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// K-major, no swizzle: [0,14) addr>>4, [16,30) LBO>>4 (K-direction core-matrix stride), [32,46) SBO>>4
// (8-row-group stride), [46,48) version = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit_to(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}

// one lane of a converged warp (the compiler keeps operands in uniform registers inside the branch)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}

// ELU(alpha = 1) without predicates: max(v, 2^(min(v,0)*log2 e) - 1)   (e^v - 1 >= v for v <= 0)
__device__ __forceinline__ float elu_bl(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.0f) * 1.4426950408889634f));
    return fmaxf(v, e - 1.0f);
}

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

