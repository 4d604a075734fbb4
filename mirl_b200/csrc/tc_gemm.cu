// Per-layer ensemble GEMMs on tcgen05 (see tc_gemm.h).
//
// One CTA computes one 128 x NT tile of one member's GEMM, fp32 in / fp32 out, through the bf16x3
// split the rollout kernel uses (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation in TMEM:
// ~2^-16 relative per product, inside the 1e-3 parity budget with two orders of margin).  Neither
// operand is pre-packed -- in training the weights change every step -- so all 256 threads read
// the fp32 operands from global memory (L2-resident at these sizes), apply the optional input
// activation, split them and write UMMA core matrices (K-major, no swizzle) into a two-stage
// shared-memory ring of 64 reduction elements; one elected thread issues the 12 MMAs of a stage
// while everybody stages the next one.  The epilogue reads the accumulator tile from TMEM
// (thread = row) and applies the mode's tail: bias + activation (forward), * swish'(z) (dX),
// + weight decay and the bias-gradient row (dW).
//
// Both operand orientations are staged without a transpose pass: "row-contiguous" operands
// (8 reduction elements adjacent in memory) are read as two 16-byte loads per thread,
// "strided" operands (reduction index is the slow one) as 8 loads that are coalesced across the
// warp because lanes own consecutive rows of the tile.
#include "tc_gemm.h"
#include "common.cuh"
#include "pe_kernels.h"
#include "tc_ptx.cuh"

namespace mb {

constexpr int kTgThreads = 256;
constexpr int kTgKC = 64;                 // reduction elements per stage = 4 UMMA k-steps
constexpr int kTgAPlane = kUmmaM * kTgKC * 2;   // bytes of one A plane (hi or lo) of a stage: 16 KB

__host__ __device__ inline int tg_ntile(int N) {      // N-tile width: <= 128 real columns, multiple of 16
    const int nsplit = (N + 127) / 128;
    const int per = (N + nsplit - 1) / nsplit;
    return (per + 15) & ~15;
}

// 8 reduction elements [rr, rr+8) of row `g` of an operand -> v (zero outside the operand).
// strided == 0: element (g, r) = src[g*ld + r];  strided == 1: element (g, r) = src[r*ld + g]
__device__ __forceinline__ void tg_load8(const float* __restrict__ src, int64_t ld, int strided, int g, int rows_valid,
                                         int ones_row, int rr, int R, float* v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (g == ones_row) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = rr + i < R ? 1.f : 0.f;
        return;
    }
    if (g >= rows_valid || rr >= R) return;
    if (!strided) {
        const float* p = src + (int64_t)g * ld + rr;
        if (rr + 8 <= R && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(p));
            const float4 y = __ldg(reinterpret_cast<const float4*>(p) + 1);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (rr + i < R) v[i] = __ldg(p + i);
        }
    } else {
        const float* p = src + (int64_t)rr * ld + g;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (rr + i < R) v[i] = __ldg(p + (int64_t)i * ld);
    }
}

// Stage `tile_rows` x 64 of an operand: items (row, 8-element k-chunk), lanes own consecutive rows.
template <int BATCH>
__device__ __forceinline__ void tg_stage(uint32_t hi_plane, uint32_t lo_plane, int tile_rows, const float* __restrict__ src,
                                         int64_t ld, int trans, int act, int row0, int rows_valid, int ones_row,
                                         int r0, int R) {
    const int items = tile_rows * (kTgKC / 8);
    const uint32_t lbo = (uint32_t)tile_rows * 16u;
    for (int base = 0; base < items; base += BATCH * kTgThreads) {
        float v[BATCH][8];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {                        // all loads of the batch in flight first
            const int it = base + q * kTgThreads + (int)threadIdx.x;
            if (it < items) {
                const int row = it % tile_rows, j = it / tile_rows;
                tg_load8(src, ld, trans, row0 + row, rows_valid, ones_row, r0 + 8 * j, R, v[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            const int it = base + q * kTgThreads + (int)threadIdx.x;
            if (it < items) {
                const int row = it % tile_rows, j = it / tile_rows;
                if (act >= 0 && row0 + row != ones_row) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[q][i] = act == MB_ACT_RELU ? fmaxf(v[q][i], 0.f) : fast_swish(v[q][i]);
                }
                const uint32_t off = (uint32_t)j * lbo + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
                split_store8(v[q], hi_plane + off, lo_plane + off);
            }
        }
    }
}

__global__ void __launch_bounds__(kTgThreads, 1)
k_tc_gemm(TgBatch gb, int nt_max) {
    extern __shared__ uint8_t tg_smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float red[kTgThreads / 32];
    const TgOp& op = gb.op[blockIdx.z / gb.nmem];
    const int zi = (int)(blockIdx.z % gb.nmem);
    const int e = gb.member[zi], slot = gb.slot0 + zi;
    const int NT = tg_ntile(op.N);
    const int m0 = blockIdx.y * kUmmaM, n0 = blockIdx.x * NT;
    if (m0 >= op.M || n0 >= op.N) return;                  // grid is sized for the largest op of the batch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t base = (smem_u32(tg_smem_raw) + 1023u) & ~1023u;
    const uint32_t b_plane = (uint32_t)nt_max * kTgKC * 2;                 // bytes of one B plane
    // stage s: [A hi | A lo | B hi | B lo]
    const uint32_t stage_bytes = 2 * kTgAPlane + 2 * b_plane;
    auto a_hi = [&](int s) { return base + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return a_hi(s) + kTgAPlane; };
    auto b_hi = [&](int s) { return a_hi(s) + 2 * kTgAPlane; };
    auto b_lo = [&](int s) { return b_hi(s) + b_plane; };
    const uint32_t bar0 = smem_u32(bars);

    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }

    const float* A = op.a + (int64_t)slot * op.a_ss;
    const float* Bm = op.b + (int64_t)slot * op.b_ss + (int64_t)e * op.b_ms;
    const int a_rows = op.a_ones_row >= 0 ? op.a_ones_row : op.M;
    const int R = op.R;
    const int nch = (R + kTgKC - 1) / kTgKC;

    auto stage = [&](int c, int s) {
        const int r0 = c * kTgKC;
        tg_stage<4>(a_hi(s), a_lo(s), kUmmaM, A, op.lda, op.a_trans, op.a_act, m0, a_rows, op.a_ones_row, r0, R);
        // tg_load8's flag means "reduction index is the strided one": that is b_trans == 0 for B'
        tg_stage<4>(b_hi(s), b_lo(s), NT, Bm, op.ldb, op.b_trans ? 0 : 1, -1, n0, op.N, -1, r0, R);
    };

    stage(0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma_idesc(NT);
    uint32_t ph0 = 0, ph1 = 0;
    for (int c = 0; c < nch; ++c) {
        const int s = c & 1;
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const int left = R - c * kTgKC;
                const int nks = left >= kTgKC ? kTgKC / 16 : (left + 15) / 16;
                const uint64_t ah = umma_desc(a_hi(s), kUmmaM * 16, 128), al = umma_desc(a_lo(s), kUmmaM * 16, 128);
                const uint64_t bh = umma_desc(b_hi(s), (uint32_t)NT * 16u, 128), bl = umma_desc(b_lo(s), (uint32_t)NT * 16u, 128);
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t ao = (uint64_t)((uint32_t)ks * (kUmmaM * 32u) >> 4);   // a k-step = 2 k-chunks
                    const uint64_t bo = (uint64_t)((uint32_t)ks * ((uint32_t)NT * 32u) >> 4);
                    tc_mma_bf16(tmem, ah + ao, bh + bo, idesc, (c | ks) ? 1u : 0u);
                    tc_mma_bf16(tmem, al + ao, bh + bo, idesc, 1u);
                    tc_mma_bf16(tmem, ah + ao, bl + bo, idesc, 1u);
                }
                tc_commit(bar0 + 8u * (uint32_t)s);
            }
            __syncwarp();
        }
        if (c + 1 < nch) {
            const int ns = s ^ 1;
            if (c >= 1) {                       // the MMAs of chunk c-1 have finished reading stage ns
                if (ns == 0) { mbar_wait_relaxed(bar0, ph0); ph0 ^= 1u; }
                else { mbar_wait_relaxed(bar0 + 8, ph1); ph1 ^= 1u; }
            }
            stage(c + 1, ns);
            fence_proxy_async();
        }
        tc_fence_before();
        __syncthreads();
    }
    {
        const int s = (nch - 1) & 1;
        mbar_wait_relaxed(bar0 + 8u * (uint32_t)s, s == 0 ? ph0 : ph1);
        tc_fence_after();
    }

    // ---- epilogue: thread = row (TMEM lane), warps w and w+4 share a lane quarter and split the columns
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float l2 = 0.f;
    for (int cb = half; cb * 16 < NT; cb += 2) {
        uint32_t acc[16];
        tc_ld16(tmem + lane_addr + (uint32_t)(cb * 16), acc);
        tc_wait_ld();
        const int nb = n0 + cb * 16;
        if (nb >= op.N) continue;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
        if (op.mode == TG_FWD) {
            if (m < op.M) {
                const float* bias = op.bias + (int64_t)e * op.bias_ms;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (nb + i < op.N) {
                        const float x = v[i] + __ldg(bias + nb + i);
                        v[i] = op.out_act < 0 ? x : (op.out_act == MB_ACT_RELU ? fmaxf(x, 0.f) : fast_swish(x));
                    }
                }
            }
        } else if (op.mode == TG_DX) {
            if (m < op.M) {
                const float* z = op.z + (int64_t)slot * op.z_ss + (int64_t)m * op.ldz + nb;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (nb + i < op.N) {
                        const float zz = __ldg(z + i);
                        const float sg = rcp_ftz(1.0f + ex2_ftz(zz * -1.4426950408889634f));
                        v[i] *= sg * (1.f + zz * (1.f - sg));
                    }
                }
            }
        } else {                                                           // TG_DW
            if (m < op.a_ones_row) {
                const float* w = op.w + (int64_t)e * op.w_ms + (int64_t)m * op.N + nb;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (nb + i < op.N) {
                        const float ww = __ldg(w + i);
                        v[i] += op.wd * ww;
                        l2 += 0.5f * op.wd * ww * ww;
                    }
                }
            }
        }
        float* dst = nullptr;
        if (op.mode == TG_DW && m == op.a_ones_row) dst = op.out_ones + (int64_t)e * op.out_ones_ms + nb;
        else if (m < (op.mode == TG_DW ? op.a_ones_row : op.M))
            dst = op.out + (int64_t)slot * op.out_ss + (int64_t)e * op.out_ms + (int64_t)m * op.ldo + nb;
        if (dst != nullptr) {
            if (nb + 16 <= op.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (nb + i < op.N) dst[i] = v[i];
            }
        }
    }
    if (op.mode == TG_DW && op.l2_acc != nullptr) {
        for (int o = 16; o; o >>= 1) l2 += __shfl_xor_sync(0xffffffffu, l2, o);
        if (lane == 0) red[warp] = l2;
    }
    tc_fence_before();
    __syncthreads();
    if (op.mode == TG_DW && op.l2_acc != nullptr && threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kTgThreads / 32; ++w) t += red[w];
        if (t != 0.f) atomicAdd(op.l2_acc, t);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int launch_tc_gemm(const TgBatch& b, cudaStream_t s) {
    int nt_max = 16, gx = 1, gy = 1;
    for (int i = 0; i < b.nops; ++i) {
        const TgOp& o = b.op[i];
        const int nt = tg_ntile(o.N);
        nt_max = nt > nt_max ? nt : nt_max;
        const int tx = (o.N + nt - 1) / nt, ty = (o.M + kUmmaM - 1) / kUmmaM;
        gx = tx > gx ? tx : gx;
        gy = ty > gy ? ty : gy;
    }
    const size_t smem = 2 * (2 * (size_t)kTgAPlane + 2 * (size_t)nt_max * kTgKC * 2) + 1024;
    static size_t configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("k_tc_gemm smem %zu: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        configured = smem;
    }
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(b.nops * b.nmem));
    k_tc_gemm<<<grid, kTgThreads, smem, s>>>(b, nt_max);
    MB_CUDA_LAUNCH_CHECK("k_tc_gemm");
    return MB_OK;
}

}  // namespace mb
