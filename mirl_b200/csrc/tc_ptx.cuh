// tcgen05 / mbarrier / bulk-copy PTX wrappers and the bf16 hi/lo split shared by the tensor-core
// kernels (pe_tc.cu: fused rollout MLP; tc_gemm.cu: per-layer ensemble GEMMs for training and the
// critic targets).  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "common.cuh"

namespace mb {

constexpr int kUmmaM = 128;        // rows of every UMMA tile (cta_group::1)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trapped launch (reported by the next CUDA
// call), never as a hung GPU.  try_wait itself suspends the thread for a HW time slice.
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
    printf("mirl_b200: mbarrier timeout block %d thread %d bar +%u parity %u\n", (int)blockIdx.x,
           (int)threadIdx.x, bar, parity);
    __trap();
}
// latency-critical wait (MMA issuer): tight poll
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) mbar_timeout(bar, parity);
}
// every other waiter backs off between polls: a spinning warp otherwise burns issue slots of
// the SM sub-partition it shares with the warps that are doing the work it is waiting for
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        __nanosleep(40);
        if (spin > (1u << 24)) mbar_timeout(bar, parity);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// A operand from tensor memory (".ts" form): [a_tmem] = 128 lanes x 8 columns per k-step, two bf16
// per 32-bit column
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tc_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) addr>>4 | [16,30) LBO>>4 (K-adjacent core matrices) | [32,46) SBO>>4 (M/N-adjacent)
// | [46,48) version = 1 | [61,64) layout = 0.  Advancing the start address by `bytes` is an
// add of bytes>>4 on the low word (no carry out of the 14-bit field inside 227 KB).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128
__host__ __device__ inline uint32_t umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);
}

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// z * sigmoid(z) = z / (1 + 2^(-z*log2e)); MUFU.EX2 + MUFU.RCP are ~2 ulp, far inside the budget
__device__ __forceinline__ float fast_swish(float z) {
    return z * rcp_ftz(1.0f + ex2_ftz(z * -1.4426950408889634f));
}
// ---- packed fp32 pairs (sm_100 FMUL2 / FADD2 / FFMA2): one issue slot for two lanes of work.  The
// epilogue is issue/latency bound (moving MUFU work to the FMA pipe made it slower), so fewer
// instructions per element is what pays.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// {hi(b) : hi(a)} packed bf16x2 (a in the low half = lower address)
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// split 8 fp32 values into bf16 hi / lo and store each as one 16-byte row of a core matrix
__device__ __forceinline__ void split_store8(const float* v, uint32_t hi_addr, uint32_t lo_addr) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
        const float ha = __uint_as_float(h[i] << 16), hb = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = cvt_bf16x2(v[2 * i] - ha, v[2 * i + 1] - hb);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}
// split 16 fp32 values into packed bf16 hi (8 words, to registers) and lo (two 16-byte core-matrix rows)
__device__ __forceinline__ void split16_hi_regs_lo_smem(const float* v, uint32_t (&h)[8], uint32_t lo_addr0, uint32_t lo_addr1) {
    uint32_t l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = cvt_bf16x2(v[2 * i], v[2 * i + 1]);
        const float ha = __uint_as_float(h[i] << 16), hb = __uint_as_float(h[i] & 0xffff0000u);
        l[i] = cvt_bf16x2(v[2 * i] - ha, v[2 * i + 1] - hb);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr1), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
}
// softplus / clamp with MUFU-based exp/log for the head (absolute error ~1e-7 on a log-std)
__device__ __forceinline__ float fast_softplus(float x) { return x > 20.0f ? x : __logf(1.0f + __expf(x)); }
__device__ __forceinline__ float fast_clamp_logstd(float raw, float mx, float mn) {
    const float ls = mx - fast_softplus(mx - raw);
    return mn + fast_softplus(ls - mn);
}


}  // namespace mb
