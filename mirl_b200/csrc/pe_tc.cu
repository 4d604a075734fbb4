// tcgen05 fused rollout step (sm_100a): the whole PEModel.step of one 128-row tile in one CTA.
//
//   reference ops replaced: base_model.py:54-76 (normalise, de-normalise, next_obs, done) +
//   pe_model.py:153-189 (MLP stack, soft log-std clamp, exp, member gather, eps*std+mean, split).
//
// Arithmetic.  Each layer GEMM runs on the 5th-gen tensor cores as THREE bf16 MMAs with fp32
// accumulation in TMEM ("bf16x3" split precision):
//        x = x_hi + x_lo,  W = W_hi + W_lo  (bf16 each)     D += x_hi*W_hi + x_lo*W_hi + x_hi*W_lo
// which is fp32-class (norm-wise error ~4e-6 vs the fp64 reference; a single bf16 or TF32 MMA
// misses the 1e-3 element-wise bar).
//
// Data movement.  Rows are bucketed by drawn member, so a tile needs ONE member's weights.  The
// weights are pre-split and pre-arranged (k_tc_pack) in exactly the UMMA shared-memory layout
// (K-major, no swizzle, 8x8 core matrices), one contiguous "stage" per 16-deep k-step, so the
// producer warp streams them from L2 with plain 1-D bulk async copies (cp.async.bulk ->
// UBLKCP) into a ring; no tensor maps.  Activations never leave the SM: the epilogue warps read
// the accumulator from TMEM (tcgen05.ld), add bias, apply swish, split to bf16 hi/lo and write
// the next layer's A operand straight into shared memory in UMMA layout.
//
// Pipeline (320 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-7  epilogue: warp w owns TMEM lanes 32*(w%4).. and the 16-column blocks of parity w/4
//   warp  8    MMA issuer (one elected lane) + TMEM allocation
//   warp  9    weight producer (one elected lane)
// The epilogue of layer l publishes its output per 16-column block (= one k-step of layer l+1)
// on a_ready[ks]; the MMA warp starts layer l+1 on block 0 while the epilogue is still working
// on the later blocks, accumulating into the OTHER of two TMEM accumulators.  The head of tile t
// (4 warps) overlaps the input staging of tile t+1 (other 4 warps) and its first-layer MMAs.
#include <cuda_bf16.h>
#include "common.cuh"
#include "pe_kernels.h"

namespace mb {

constexpr int kTcThreads = 320;
constexpr int kTcTM = 128;            // rows per tile (UMMA M)
constexpr int kTcMaxKs = 16;          // k-steps per layer supported (HP <= 256)
constexpr int kTcMaxStages = 6;
constexpr int kTcSmemLimit = 227 * 1024;

// ---- shapes derived from the model ---------------------------------------------------------
struct TcShape {
    int L;        // linear layers
    int K0P;      // padded input width (multiple of 16)
    int HP;       // padded hidden width (multiple of 16, <= 256)
    int NOP;      // output tile width: 64 (two 32-column halves: mean | raw log-std)
    int nks0;     // k-steps of layer 0
    int nksh;     // k-steps of the other layers
    int stages;   // weight ring depth
    int slot;     // bytes per ring slot
    int64_t member_bytes;   // pack bytes per member
    bool ok;
};

__host__ __device__ inline int tc_layer_np(const TcShape& s, int l) { return l == s.L - 1 ? s.NOP : s.HP; }
__host__ __device__ inline int tc_layer_nks(const TcShape& s, int l) { return l == 0 ? s.nks0 : s.nksh; }
__host__ __device__ inline int tc_stage_bytes(const TcShape& s, int l) { return 64 * tc_layer_np(s, l); }
__host__ __device__ inline int64_t tc_layer_off(const TcShape& s, int l) {
    int64_t off = 0;
    for (int i = 0; i < l; ++i) off += (int64_t)tc_layer_nks(s, i) * tc_stage_bytes(s, i);
    return off;
}

static int ru16(int x) { return (x + 15) & ~15; }

static size_t tc_smem_bytes(const TcShape& s, int O) {
    size_t a = (size_t)2 * kTcTM * s.HP * 2;          // A hi + lo
    size_t ring = (size_t)s.stages * s.slot;
    size_t misc = 2 * (size_t)kTcTM * O * 4 + 2 * kTcTM * 4 + 1024;
    return a + ring + misc + 1024;                    // + alignment slack
}

static TcShape tc_shape(const mb_mlp_desc& d) {
    TcShape s;
    s.ok = false;
    s.L = d.num_layers;
    if (s.L < 2 || d.activation != MB_ACT_SWISH) return s;
    const int H = d.sizes[1];
    for (int l = 1; l < s.L; ++l)
        if (d.sizes[l] != H) return s;
    s.K0P = ru16(d.sizes[0]);
    s.HP = ru16(H);
    s.NOP = 64;                       // mean at columns [0,D), raw log-std at [32,32+D)
    if (s.HP > 256 || s.K0P > 32 || s.HP < 64 || d.sizes[s.L] / 2 > 32) return s;
    s.nks0 = s.K0P / 16;
    s.nksh = s.HP / 16;
    s.slot = 64 * s.HP;
    s.member_bytes = tc_layer_off(s, s.L);
    s.stages = kTcMaxStages;
    const int O = d.sizes[s.L] / 2 - 1;
    while (s.stages > 2 && tc_smem_bytes(s, O) > (size_t)kTcSmemLimit) --s.stages;
    if (tc_smem_bytes(s, O) > (size_t)kTcSmemLimit) return s;
    s.ok = true;
    return s;
}

size_t tc_pack_bytes(const mb_mlp_desc& d) {
    TcShape s = tc_shape(d);
    return s.ok ? (size_t)s.member_bytes * d.ensemble_size : 0;
}

// ---- weight pack: fp32 [E][K][N] -> per (member, layer, k-step): [hi | lo] blocks, each
//      [2 k-chunks][NP/8][8 n][8 k] bf16  (UMMA K-major, no swizzle: LBO = NP*16 B, SBO = 128 B)
__global__ void k_tc_pack(mb_mlp_desc d, TcShape s, const float* __restrict__ params,
                          __nv_bfloat16* __restrict__ pack) {
    const int e = blockIdx.y, l = blockIdx.z;
    const int K = d.sizes[l], N = d.sizes[l + 1];
    const int NP = tc_layer_np(s, l), nks = tc_layer_nks(s, l);
    const float* W = params + w_offset(d, l) + (size_t)e * K * N;
    __nv_bfloat16* out = pack + ((int64_t)e * s.member_bytes + tc_layer_off(s, l)) / 2;
    const int per_stage = 16 * NP;    // elements per hi (or lo) block
    const int total = nks * per_stage;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ks = i / per_stage, r = i - ks * per_stage;
        const int chunk = r / (NP * 8), r2 = r - chunk * (NP * 8);
        const int n = r2 >> 3, kk = r2 & 7;             // [n/8][n%8][k] == n*8 + k
        const int k = ks * 16 + chunk * 8 + kk;
        int nsrc = n;                                    // source column in W[K][N]
        if (l == s.L - 1) {                              // head layout: mean | pad | raw log-std | pad
            const int Dh = N / 2, half = n >> 5, j = n & 31;
            nsrc = j < Dh ? half * Dh + j : N;
        }
        const float w = (k < K && nsrc < N) ? W[(size_t)k * N + nsrc] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        __nv_bfloat16* st = out + (size_t)ks * 2 * per_stage;
        st[r] = hi;
        st[per_stage + r] = lo;
    }
}

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 26)) {
            printf("mirl_b200: mbarrier timeout block %d thread %d bar +%u parity %u\n", (int)blockIdx.x,
                   (int)threadIdx.x, bar, parity);
            __trap();
        }
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) addr>>4 | [16,30) LBO>>4 (K-adjacent core matrices) | [32,46) SBO>>4 (M/N-adjacent)
// | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128
__host__ __device__ inline uint32_t umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTcTM >> 4) << 24);
}

__device__ __forceinline__ float fast_swish(float z) {
    // z * sigmoid(z); ex2.approx + rcp.approx are ~2 ulp, far inside the bf16x3 budget
    return __fdividef(z, 1.0f + __expf(-z));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// split 8 fp32 values into bf16 hi / lo and store each as one 16-byte chunk row
__device__ __forceinline__ void split_store8(const float* v, uint32_t hi_addr, uint32_t lo_addr) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = pack_bf16x2(v[2 * i] - __bfloat162float(h0), v[2 * i + 1] - __bfloat162float(h1));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

// ---- the kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
k_pe_rollout_tc(PeDev m, TcShape s, const uint8_t* __restrict__ pack, const float* __restrict__ obs,
                const float* __restrict__ act, const float* __restrict__ noise,
                const int* __restrict__ perm, const int4* __restrict__ tiles,
                const int* __restrict__ ntiles_p, int done_kind, float* __restrict__ next_obs,
                float* __restrict__ reward, float* __restrict__ done) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const int O = m.O, A = m.A, D = O + 1, L = s.L;
    const uint32_t a_bytes = (uint32_t)kTcTM * s.HP * 2;          // one of hi / lo
    const uint32_t sA_hi = smem_base;
    const uint32_t sA_lo = sA_hi + a_bytes;
    const uint32_t sRing = sA_lo + a_bytes;
    uint8_t* misc = smem + 2 * a_bytes + (size_t)s.stages * s.slot;
    float* raw = reinterpret_cast<float*>(misc);                  // [2][TM][O]
    int* rowid = reinterpret_cast<int*>(raw + 2 * kTcTM * O);     // [2][TM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(rowid + 2 * kTcTM);
    const uint32_t bar0 = smem_u32(bars);
    auto w_full = [&](int i) { return bar0 + 8u * i; };
    auto w_empty = [&](int i) { return bar0 + 8u * (kTcMaxStages + i); };
    auto a_ready = [&](int i) { return bar0 + 8u * (2 * kTcMaxStages + i); };
    auto d_full = [&](int i) { return bar0 + 8u * (2 * kTcMaxStages + kTcMaxKs + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcMaxStages + kTcMaxKs + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = *ntiles_p;
    const int n_my = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < s.stages; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
        for (int i = 0; i < kTcMaxKs; ++i) mbar_init(a_ready(i), 4);
        mbar_init(d_full(0), 1);
        mbar_init(d_full(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const float* nrm = m.norm;

    if (warp < 8) {
        // ================================ EPILOGUE WARPS ========================================
        const int q = warp & 3, p = warp >> 2;
        const int row = q * 32 + lane;                            // tile row == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t row_off = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;

        // Stage the normalised inputs of local tile i as layer-0 A operand (thread = row).
        auto stage_tile = [&](int i) {
            const int4 ti = tiles[blockIdx.x + i * gridDim.x];
            const int buf = i & 1;
            float x[32];
            const int r = row < ti.z ? perm[ti.y + row] : -1;
            const float* op = obs + (size_t)(r < 0 ? 0 : r) * O;
            const float* ap = act + (size_t)(r < 0 ? 0 : r) * A;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float v = 0.f;
                if (r >= 0) {
                    if (j < O) {
                        const float t = op[j];
                        raw[(buf * kTcTM + row) * O + j] = t;
                        v = (t - nrm[j]) / (nrm[O + j] + kNormEps);
                    } else if (j < O + A) {
                        v = (ap[j - O] - nrm[2 * O + (j - O)]) / (nrm[2 * O + A + (j - O)] + kNormEps);
                    }
                }
                x[j] = v;
            }
            rowid[buf * kTcTM + row] = r;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c * 8 < s.K0P)
                    split_store8(x + 8 * c, sA_hi + c * 2048u + row_off, sA_lo + c * 2048u + row_off);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
                for (int ks = 0; ks < s.nks0; ++ks) mbar_arrive(a_ready(ks));
        };

        if (p == 1 && n_my > 0) stage_tile(0);
        for (int i = 0; i < n_my; ++i) {
            const int4 ti = tiles[blockIdx.x + i * gridDim.x];
            const int e = ti.x;
            for (int l = 0; l + 1 < L; ++l) {
                const int g = i * L + l;
                mbar_wait(d_full(g & 1), (g >> 1) & 1);
                tc_fence_after();
                const float* bias = m.params + b_offset(m.mlp, l) + (size_t)e * m.mlp.sizes[l + 1];
                const int nvalid = m.mlp.sizes[l + 1];
                const uint32_t dcol = tmem + lane_addr + (uint32_t)((g & 1) * 256);
                for (int ks = p; ks < s.nksh; ks += 2) {
                    uint32_t rr[16];
                    tc_ld16(dcol + ks * 16, rr);
                    tc_wait_ld();
                    float v[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int col = ks * 16 + c;
                        const float b = col < nvalid ? __ldg(bias + col) : 0.f;
                        v[c] = fast_swish(__uint_as_float(rr[c]) + b);
                    }
                    split_store8(v, sA_hi + (2 * ks) * 2048u + row_off, sA_lo + (2 * ks) * 2048u + row_off);
                    split_store8(v + 8, sA_hi + (2 * ks + 1) * 2048u + row_off, sA_lo + (2 * ks + 1) * 2048u + row_off);
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_ready(ks));
                }
            }
            // ---- last layer: head (p == 0) overlapped with next tile's input staging (p == 1)
            const int g = i * L + L - 1;
            mbar_wait(d_full(g & 1), (g >> 1) & 1);
            tc_fence_after();
            if (p == 1) {
                if (i + 1 < n_my) stage_tile(i + 1);
            } else {
                const int buf = i & 1;
                const int r = rowid[buf * kTcTM + row];
                const uint32_t dcol = tmem + lane_addr + (uint32_t)((g & 1) * 256);
                float mu_[32], rl_[32];
                {
                    uint32_t rr[16];
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        tc_ld16(dcol + blk * 16, rr);
                        tc_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) mu_[blk * 16 + c] = __uint_as_float(rr[c]);
                        tc_ld16(dcol + 32 + blk * 16, rr);
                        tc_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) rl_[blk * 16 + c] = __uint_as_float(rr[c]);
                    }
                }
                tc_fence_before();
                if (r >= 0) {
                    const float* bias = m.params + b_offset(m.mlp, L - 1) + (size_t)e * 2 * D;
                    const float* mxp = m.params + logstd_offset(m.mlp, 0) + (size_t)e * D;
                    const float* mnp = m.params + logstd_offset(m.mlp, 1) + (size_t)e * D;
                    const float* dmean = nrm + 2 * O + 2 * A;
                    const float* dstd = dmean + O;
                    const float* eps = noise + (size_t)r * D;
                    float* ro = raw + (buf * kTcTM + row) * O;      // raw obs in, next obs out
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < D) {
                            const float mu = mu_[j] + __ldg(bias + j);
                            const float rawls = rl_[j] + __ldg(bias + D + j);
                            const float ls = clamp_logstd(rawls, __ldg(mxp + j), __ldg(mnp + j));
                            const float sv = eps[j] * expf(ls) + mu;
                            if (j < O) {
                                const float v = ro[j] + (sv * (dstd[j] + kNormEps) + dmean[j]);
                                ro[j] = v;
                                next_obs[(size_t)r * O + j] = v;
                            } else {
                                reward[r] = sv;
                            }
                        }
                    }
                    done[r] = done_fn(done_kind, ro, O);
                }
            }
        }
    } else if (warp == 8) {
        // ================================ MMA ISSUER ============================================
        if (lane == 0) {
            uint32_t aphase = 0;           // bit ks = parity to wait for on a_ready[ks]
            int stage = 0;
            uint32_t wphase = 0;
            for (int i = 0; i < n_my; ++i) {
                for (int l = 0; l < L; ++l) {
                    const int g = i * L + l;
                    const int nks = tc_layer_nks(s, l), NP = tc_layer_np(s, l);
                    const uint32_t idesc = umma_idesc(NP);
                    const uint32_t d_tmem = tmem + (uint32_t)((g & 1) * 256);
                    const uint32_t lbo_b = (uint32_t)NP * 16u;
                    for (int ks = 0; ks < nks; ++ks) {
                        mbar_wait(a_ready(ks), (aphase >> ks) & 1u);
                        aphase ^= 1u << ks;
                        mbar_wait(w_full(stage), wphase);
                        tc_fence_after();
                        const uint32_t wb = sRing + (uint32_t)stage * s.slot;
                        const uint64_t a_hi = umma_desc(sA_hi + ks * 4096u, 2048u, 128u);
                        const uint64_t a_lo = umma_desc(sA_lo + ks * 4096u, 2048u, 128u);
                        const uint64_t b_hi = umma_desc(wb, lbo_b, 128u);
                        const uint64_t b_lo = umma_desc(wb + 32u * NP, lbo_b, 128u);
                        tc_mma_bf16(d_tmem, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
                        tc_mma_bf16(d_tmem, a_lo, b_hi, idesc, 1u);
                        tc_mma_bf16(d_tmem, a_hi, b_lo, idesc, 1u);
                        tc_commit(w_empty(stage));
                        if (++stage == s.stages) { stage = 0; wphase ^= 1u; }
                    }
                    tc_commit(d_full(g & 1));
                }
            }
        }
    } else {
        // ================================ WEIGHT PRODUCER ========================================
        if (lane == 0) {
            int stage = 0;
            uint32_t wphase = 0;
            for (int i = 0; i < n_my; ++i) {
                const int e = tiles[blockIdx.x + i * gridDim.x].x;
                const uint8_t* mp = pack + (size_t)e * s.member_bytes;
                for (int l = 0; l < L; ++l) {
                    const int nks = tc_layer_nks(s, l);
                    const uint32_t bytes = (uint32_t)tc_stage_bytes(s, l);
                    const uint8_t* lp = mp + tc_layer_off(s, l);
                    for (int ks = 0; ks < nks; ++ks) {
                        mbar_wait(w_empty(stage), wphase ^ 1u);
                        mbar_expect_tx(w_full(stage), bytes);
                        bulk_g2s(sRing + (uint32_t)stage * s.slot, lp + (size_t)ks * bytes, bytes, w_full(stage));
                        if (++stage == s.stages) { stage = 0; wphase ^= 1u; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---- host launchers -----------------------------------------------------------------------------
int launch_tc_pack(const PeDev& m, void* pack, cudaStream_t st) {
    TcShape s = tc_shape(m.mlp);
    if (!s.ok) { set_error("tcgen05 rollout kernel does not support this model shape"); return MB_EUNSUPPORTED; }
    dim3 grid(16, (unsigned)m.mlp.ensemble_size, (unsigned)s.L);
    k_tc_pack<<<grid, 256, 0, st>>>(m.mlp, s, m.params, reinterpret_cast<__nv_bfloat16*>(pack));
    MB_CUDA_LAUNCH_CHECK("k_tc_pack");
    return MB_OK;
}

int launch_pe_rollout_tc(const PeDev& m, const void* pack, const float* obs, const float* act,
                         const uint8_t* idx, const float* noise, int64_t B, int done_kind,
                         float* next_obs, float* reward, float* done, void* ws, cudaStream_t st) {
    TcShape s = tc_shape(m.mlp);
    if (!s.ok) { set_error("tcgen05 rollout kernel does not support this model shape"); return MB_EUNSUPPORTED; }
    if (m.O > 63 || m.O + m.A > 32) { set_error("tcgen05 rollout kernel: obs+act must be <= 32"); return MB_EUNSUPPORTED; }
    BucketPlan p = bucket_rows(idx, B, m.mlp.ensemble_size, kTcTM, ws, st);
    MB_CUDA_LAUNCH_CHECK("bucket_rows");
    const size_t smem = tc_smem_bytes(s, m.O);
    static size_t configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_pe_rollout_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("k_pe_rollout_tc smem %zu: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        configured = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.max_tiles < sms ? p.max_tiles : sms;
    prof_begin(st);
    k_pe_rollout_tc<<<grid, kTcThreads, smem, st>>>(m, s, reinterpret_cast<const uint8_t*>(pack), obs, act, noise,
                                                    p.perm, p.tiles, p.ntiles, done_kind, next_obs, reward, done);
    prof_end(st);
    MB_CUDA_LAUNCH_CHECK("k_pe_rollout_tc");
    return MB_OK;
}

}  // namespace mb
