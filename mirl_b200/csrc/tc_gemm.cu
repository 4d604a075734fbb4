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
#include <stdlib.h>

namespace mb {

constexpr int kTgThreads = 512;            // 4 warps per SM sub-partition: the staging code is latency bound
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

// Stage a tile_rows x 64 block of an operand.  Every thread owns NI items (row, 8-element k-chunk)
// whose coordinates do not depend on the stage, with lanes on consecutive rows: all global loads
// of the thread are issued before the first conversion.
struct TgItem { int row, j; };
template <int NI>
__device__ __forceinline__ void tg_fetch(float (&v)[NI][8], const TgItem (&it)[NI], const float* __restrict__ src,
                                         int64_t ld, int strided, int row0, int rows_valid, int ones_row, int r0, int R) {
#pragma unroll
    for (int q = 0; q < NI; ++q)
        if (it[q].row >= 0) tg_load8(src, ld, strided, row0 + it[q].row, rows_valid, ones_row, r0 + 8 * it[q].j, R, v[q]);
}
template <int NI>
__device__ __forceinline__ void tg_convert(float (&v)[NI][8], uint32_t hi_plane, uint32_t lo_plane, uint32_t lbo,
                                           const TgItem (&it)[NI], int act, int row0, int ones_row) {
#pragma unroll
    for (int q = 0; q < NI; ++q) {
        if (it[q].row < 0) continue;
        if (act >= 0 && row0 + it[q].row != ones_row) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[q][i] = act == MB_ACT_RELU ? fmaxf(v[q][i], 0.f) : fast_swish(v[q][i]);
        }
        const uint32_t off = (uint32_t)it[q].j * lbo + (uint32_t)(it[q].row >> 3) * 128u + (uint32_t)(it[q].row & 7) * 16u;
        split_store8(v[q], hi_plane + off, lo_plane + off);
    }
}

__global__ void __launch_bounds__(kTgThreads, 1)
k_tc_gemm(TgBatch gb, int nt_max) {
    extern __shared__ uint8_t tg_smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float red[kTgThreads / 32];
    // this CTA's descriptor -> shared memory once (indexing the kernel-parameter array with a
    // run-time index makes every later field access a slow indexed constant load)
    __shared__ TgOp sop;
    pdl_trigger();
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&gb.op[blockIdx.z / gb.nmem]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sop);
        for (int i = threadIdx.x; i < (int)(sizeof(TgOp) / 4); i += kTgThreads) dst[i] = src[i];
    }
    __syncthreads();
    const TgOp& op = sop;
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

    // item coordinates (fixed for the whole kernel): A 128 x 8 items, B NT x 8 items (NT <= 128)
    constexpr int kItems = kUmmaM * (kTgKC / 8) / kTgThreads;         // 2
    const int a_act = op.a_act, a_strided = op.a_trans, a_ones = op.a_ones_row, b_strided = op.b_trans ? 0 : 1;
    // Two lane mappings.  Operand stored with the reduction index strided (element (g, r) at
    // src[r*ld + g]): lanes on 32 consecutive rows g, so each scalar load of the warp is one 128 B
    // line.  Reduction index contiguous: a warp covers 8 rows x 4 k-chunks -- the 4 chunks of a row
    // are 128 contiguous bytes, so a 16 B load instruction touches 8 lines instead of 32, and the
    // 8 lanes of a quarter warp write the 8 rows of one core matrix (128 contiguous bytes of
    // shared memory, no bank conflict).
    TgItem ia[kItems], ib[kItems];
#pragma unroll
    for (int q = 0; q < kItems; ++q) {
        const int it = (int)threadIdx.x + q * kTgThreads;
        const int row8 = 8 * warp + (lane & 7), j8 = 4 * q + (lane >> 3);
        if (a_strided) { ia[q].row = it & (kUmmaM - 1); ia[q].j = it / kUmmaM; }
        else { ia[q].row = row8; ia[q].j = j8; }
        if (b_strided) { ib[q].row = it < NT * (kTgKC / 8) ? it % NT : -1; ib[q].j = it / NT; }
        else { ib[q].row = row8 < NT ? row8 : -1; ib[q].j = j8; }
    }
    const int64_t lda = op.lda, ldb = op.ldb;
    const int Nn = op.N;
    // stage = fetch (global -> registers, both operands in flight together) + convert (registers ->
    // bf16 hi/lo planes of ring slot s); `ready` runs between the two: the loads overlap whatever
    // it waits for
    auto stage = [&](int c, int s, auto&& ready) {
        const int r0 = c * kTgKC;
        float va[kItems][8], vb[kItems][8];
        tg_fetch<kItems>(va, ia, A, lda, a_strided, m0, a_rows, a_ones, r0, R);
        // "reduction index is the strided one" is b_trans == 0 for B'
        tg_fetch<kItems>(vb, ib, Bm, ldb, b_strided, n0, Nn, -1, r0, R);
        ready();
        tg_convert<kItems>(va, a_hi(s), a_lo(s), kUmmaM * 16u, ia, a_act, m0, a_ones);
        tg_convert<kItems>(vb, b_hi(s), b_lo(s), (uint32_t)NT * 16u, ib, -1, n0, -1);
    };

    pdl_wait();            // everything above touched only shared / tensor memory
    stage(0, 0, [] {});
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
            stage(c + 1, ns, [&] {
                if (c >= 1) {                   // the MMAs of chunk c-1 have finished reading stage ns
                    if (ns == 0) { mbar_wait_relaxed(bar0, ph0); ph0 ^= 1u; }
                    else { mbar_wait_relaxed(bar0 + 8, ph1); ph1 ^= 1u; }
                }
            });
            fence_proxy_async();
        }
        tc_fence_before();
        __syncthreads();
        }
    // epilogue constants and this lane's bias columns are fetched before the last MMAs are waited for
    const int mode = op.mode, M = op.M, out_act = op.out_act, ones_row = op.a_ones_row;
    const float wd = op.wd;
    const int row_lim = mode == TG_DW ? ones_row : M;                     // rows with a regular output
    const int col = lane * 4;                                             // this lane's 4 columns of the tile
    const int nb = n0 + col;
    const bool col_ok = col < NT && nb < Nn;
    const bool full = nb + 4 <= Nn;
    float l2 = 0.f;
    float4 cb4 = make_float4(0.f, 0.f, 0.f, 0.f);                          // FWD: bias of these columns
    if (mode == TG_FWD && col_ok) {
        const float* p = op.bias + (int64_t)e * op.bias_ms + nb;
        if (full && (reinterpret_cast<uintptr_t>(p) & 15) == 0) cb4 = __ldg(reinterpret_cast<const float4*>(p));
        else { cb4.x = __ldg(p); if (nb + 1 < Nn) cb4.y = __ldg(p + 1); if (nb + 2 < Nn) cb4.z = __ldg(p + 2); if (nb + 3 < Nn) cb4.w = __ldg(p + 3); }
    }
    {
        const int s = (nch - 1) & 1;
        mbar_wait_relaxed(bar0 + 8u * (uint32_t)s, s == 0 ? ph0 : ph1);
        tc_fence_after();
    }

    // ---- epilogue.  Phase 1: accumulators TMEM -> registers (thread = row) -> shared memory (the
    // operand ring is free now) as a [128][NT + 4] fp32 tile.  Phase 2: warps walk the rows with
    // lanes along the columns, so the per-element operands (bias / z / w) and the result move as
    // coalesced 16-byte accesses; the tail op of the mode is applied here.
    const int LDT = NT + 4;                                                // floats; 16-byte rows, conflict-free
    const uint32_t tile = base;
    {
        const int q = warp & 3, grp = warp >> 2;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t rowp = tile + (uint32_t)((q * 32 + lane) * LDT) * 4u;
        for (int cb = grp; cb * 16 < NT; cb += kTgThreads / 128) {
            uint32_t acc[16];
            tc_ld16(tmem + lane_addr + (uint32_t)(cb * 16), acc);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (uint32_t)(cb * 16 + 4 * i) * 4u),
                             "r"(acc[4 * i]), "r"(acc[4 * i + 1]), "r"(acc[4 * i + 2]), "r"(acc[4 * i + 3]) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    auto ld4 = [&](const float* p) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full && (reinterpret_cast<uintptr_t>(p) & 15) == 0) x = __ldg(reinterpret_cast<const float4*>(p));
        else { x.x = __ldg(p); if (nb + 1 < Nn) x.y = __ldg(p + 1); if (nb + 2 < Nn) x.z = __ldg(p + 2); if (nb + 3 < Nn) x.w = __ldg(p + 3); }
        return x;
    };
    const float* zb = (mode == TG_DX && op.z != nullptr) ? op.z + (int64_t)slot * op.z_ss : nullptr;   // no z: plain A B^T
    const bool z_relu = op.z_act == MB_ACT_RELU;
    const float* wb = mode == TG_DW ? op.w + (int64_t)e * op.w_ms : nullptr;
    float* ob = op.out + (int64_t)slot * op.out_ss + (int64_t)e * op.out_ms;
    const int64_t ldz = op.ldz, ldo = op.ldo;
    constexpr int kRU = 8;                                                 // rows in flight per warp (128 rows / 16 warps)
    for (int rb = warp * kRU; rb < kUmmaM; rb += (kTgThreads / 32) * kRU) {
        float4 c[kRU], v[kRU];
        bool live[kRU], plain[kRU];
#pragma unroll
        for (int u = 0; u < kRU; ++u) {                                    // all loads first
            const int r = rb + u, m = m0 + r;
            plain[u] = mode == TG_DW && m == ones_row;                     // bias-gradient row: no tail op
            live[u] = (m < row_lim || plain[u]) && col_ok;
            c[u] = cb4;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!live[u]) continue;
            if (mode == TG_DX) { if (zb != nullptr) c[u] = ld4(zb + (int64_t)m * ldz + nb); }
            else if (mode == TG_DW && !plain[u]) c[u] = ld4(wb + (int64_t)m * Nn + nb);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                         : "r"(tile + (uint32_t)(r * LDT + col) * 4u));
        }
#pragma unroll
        for (int u = 0; u < kRU; ++u) {
            if (!live[u]) continue;
            const int m = m0 + rb + u;
            float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            const float cc[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
            if (mode == TG_FWD) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = vv[i] + cc[i];
                    vv[i] = out_act < 0 ? x : (out_act == MB_ACT_RELU ? fmaxf(x, 0.f) : fast_swish(x));
                }
            } else if (mode == TG_DX) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (zb == nullptr) continue;
                    if (z_relu) { vv[i] = cc[i] > 0.f ? vv[i] : 0.f; continue; }
                    const float sg = rcp_ftz(1.0f + ex2_ftz(cc[i] * -1.4426950408889634f));
                    vv[i] *= sg * (1.f + cc[i] * (1.f - sg));
                }
            } else if (!plain[u]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {                              // c = 0 beyond N: no contribution
                    vv[i] += wd * cc[i];
                    l2 += 0.5f * wd * cc[i] * cc[i];
                }
            }
            float* dst = (plain[u] ? op.out_ones + (int64_t)e * op.out_ones_ms : ob + (int64_t)m * ldo) + nb;
            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                *reinterpret_cast<float4*>(dst) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (nb + i < Nn) dst[i] = vv[i];
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
    static size_t configured[64] = {0};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev < 0 || dev >= 64 ? 0 : dev;
    if (configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("k_tc_gemm smem %zu: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        configured[dev] = smem;
    }
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(b.nops * b.nmem));
    launch_pdl(k_tc_gemm, grid, dim3(kTgThreads), smem, s, b, nt_max);
    MB_CUDA_LAUNCH_CHECK("k_tc_gemm");
    return MB_OK;
}

}  // namespace mb
