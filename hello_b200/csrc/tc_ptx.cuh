// Thin inline-PTX wrappers for the Blackwell (sm_100a) primitives the fused read convolver uses:
// mbarrier, bulk async copy (TMA unit, UBLKCP), tcgen05 MMA / TMEM alloc / TMEM load-store / commit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace hello {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait until the barrier phase with the given parity completes.  try_wait carries a suspend-time hint, so a waiting
// warp sleeps in hardware until the phase flips instead of spinning through issue slots the working warps need
// (a bare try_wait returned every ~130 cycles here and the wait loops executed as many instructions as the
// epilogues).  The loop lives inside the asm block so the compiler keeps treating the surrounding control flow as
// warp-uniform (needed for the uniform-datapath MMA issue code).  A wait that makes no progress for ~2 s (4e9 SM
// clocks) is a protocol bug: trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t.reg .u64 t0, t1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "mov.u64 t0, %%clock64;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "mov.u64 t1, %%clock64;\n\t"
        "sub.u64 t1, t1, t0;\n\t"
        "setp.lt.u64 P1, t1, 4000000000;\n\t"
        "@P1 bra LAB_WAIT;\n\t"
        "trap;\n\t"
        "DONE:\n\t}" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- bulk copy
// global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// hint: pull [p, p+bytes) into L2 (16-byte aligned address and size)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// All previously issued tcgen05.mma of the issuing thread complete -> one arrival on the mbarrier.  Executed by a
// converged warp; one elected lane issues (the same lane elect.sync picks for the MMAs).
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes stored contiguously
// (128 B); `lbo` = byte distance between core matrices adjacent in K, `sbo` = between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);   // bits [46,48) = 0b01: Blackwell descriptor version
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Instruction descriptor for kind::f16 with bf16 operands (both K-major), fp32 accumulate, M = 128.
__host__ __device__ constexpr uint32_t idesc_bf16_m128(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// High word of a K-major / no-swizzle descriptor with SBO = 128 B (8-row groups contiguous) + descriptor version 1.
constexpr uint32_t DESC_HI_SBO128 = (128u >> 4) | (1u << 14);
// Low word: start address and LBO, both in 16-byte units.  Stepping the operand by `x` bytes adds x>>4 to it.
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) {
    return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
}

// D[tmem] (+)= A[smem] * B[smem].  Executed by a converged warp with warp-uniform operands; one elected lane
// issues on behalf of the CTA (SASS: UTCHMMA).
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint32_t a_desc_lo, uint32_t b_desc_lo, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_desc_lo), "r"(b_desc_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI_SBO128)
        : "memory");
}

// TMEM -> registers: each thread reads N consecutive 32-bit columns of its own lane (SASS: LDTM)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v) {
    static_assert(N == 16 || N == 32, "column block of 16 or 32");
    if (N == 16) tmem_ld16(taddr, v); else tmem_ld32(taddr, v);
}

// registers -> TMEM (SASS: STTM)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
        : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Packed fp32x2 add / subtract (Blackwell FADD2): two independent round-to-nearest results, one issue slot.
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ta, tb;\n\tmov.b64 ta, {%0, %1};\n\tmov.b64 tb, {%2, %3};\n\t"
        "add.rn.f32x2 ta, ta, tb;\n\tmov.b64 {%0, %1}, ta;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void sub2(float& a0, float& a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ta, tb;\n\tmov.b64 ta, {%0, %1};\n\tmov.b64 tb, {%2, %3};\n\t"
        "sub.rn.f32x2 ta, ta, tb;\n\tmov.b64 {%0, %1}, ta;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}

// two fp32 -> packed bf16x2 (round to nearest even); `lo` lands in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

}  // namespace ptx
}  // namespace hello
