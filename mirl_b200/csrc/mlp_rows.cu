// Row-parallel forward pass of ONE small MLP over many rows: the policy network inside the rollout loop
// (GaussianPolicy.action -> MLP, mirl/policies/gaussian_policy.py:35-47, mirl/torch_modules/mlp.py:141-168;
// 17 -> 256 -> 256 -> 12 relu at 1e5 rows per horizon step, mbpo_collector.py:48).
//
// Round 2 ran it as three per-layer launches (a SIMT layer 0 + two k_tc_gemm launches that convert fp32
// activations to bf16 on the fly and write every layer's output back to HBM): 0.36 ms per 1e5 rows, more than the
// fused rollout step it feeds.  Here the whole network is one launch:
//   * a CTA owns 64 rows (four MMA M tiles) at a time and keeps their activations in shared memory as bf16 hi/lo
//     pairs from the input to the output layer -- nothing but the input rows and the final output touches HBM;
//   * the weights stream from L2 (a bf16 hi/lo pack in 32-row chunks, one cp.async.bulk per chunk, two-stage
//     ring with full/empty mbarriers) ONCE per 64 rows: a weight fragment loaded with ldmatrix feeds four M tiles;
//   * warp-level tensor-core MMAs (mma.sync.m16n8k16, bf16x3 split: x_hi W_hi + x_lo W_hi + x_hi W_lo, fp32
//     accumulate -- the same fp32-class arithmetic as the other tensor-core kernels); the bias rides in the GEMM
//     as weight row K against a constant-1 activation column.
// Hidden widths up to 256 (4 n-tiles of 8 columns per warp, 8 warps).  The pack is rebuilt by a small kernel at
// every call (73 k weights for the policy: microseconds), so there is no cache to invalidate when the actor
// update changes the weights.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"
#include "pe_kernels.h"
#include "tc_ptx.cuh"

namespace mb {

namespace {

constexpr int kWarps = 8;
constexpr int kT = kWarps * 32;
constexpr int kIssuer = kWarps - 1;
#ifndef MB_ROWS_BLOCK
#define MB_ROWS_BLOCK 64
#endif
#ifndef MB_ROWS_STAGES
#define MB_ROWS_STAGES 2
#endif
constexpr int kBlockRows = MB_ROWS_BLOCK;     // rows per CTA pass: its M tiles share every weight fragment
constexpr int kMT = kBlockRows / 16;
constexpr int kWChunk = 32;        // weight rows per streamed chunk (two k-steps)
constexpr int kStages = MB_ROWS_STAGES;   // ring depth.  16-row chunks in a 4-stage ring (three copies in flight) measured 30 % SLOWER:
                                   // several bulk copies in flight per SM cost more than the latency they hide
constexpr int kNT = 4;             // n-tiles per warp: widths up to 8 * kWarps * kNT = 256

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {   // element 0 in the low half
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
    return d;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(a - ha, b - hb);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }


struct RowsLayer {
    int K, N;        // fan-in, fan-out
    int Kp;          // roundup16(K + 1): pack rows, row K = bias
    int Ns;          // bf16 row stride of the pack rows (>= roundup16(N), = 8 mod 16: conflict-free ldmatrix)
    int As;          // bf16 row stride of this layer's input activations (>= Kp, = 8 mod 16)
    int nt;          // ceil(N / 8)
    int64_t pack_off;    // bf16 offset of the layer's pack
};
struct RowsArgs {
    int L, act;
    int as_max, stage_bytes, act_bytes, raw_bytes;
    RowsLayer ly[MB_MAX_LAYERS];
    const float* x;           // [B][K0]
    int64_t B;
    float* out;               // [B][N_L]
    const __nv_bfloat16* pack;
};

__host__ __device__ inline int stride_rule(int n) { return (n + 15) / 16 * 16 + 8; }
__host__ __device__ inline int chunk_count(int Kp) { return (Kp + kWChunk - 1) / kWChunk; }
// pack of a layer: kWChunk-row chunks, each [hi rows x Ns][lo rows x Ns] contiguous -> a chunk is ONE bulk copy
__host__ __device__ inline int64_t pack_row_off(int k, int half, int Kp, int Ns) {
    const int c = k / kWChunk;
    const int rows = Kp - c * kWChunk < kWChunk ? Kp - c * kWChunk : kWChunk;
    return (int64_t)c * (kWChunk * Ns * 2) + (int64_t)half * rows * Ns + (int64_t)(k - c * kWChunk) * Ns;
}

// fp32 parameter blob (member `member` of the ensemble layout) -> bf16 hi/lo pack, padding included
__global__ void k_rows_pack(mb_mlp_desc d, RowsArgs a, const float* __restrict__ params, int member,
                            __nv_bfloat16* __restrict__ pack) {
    pdl_trigger();
    pdl_wait();
    const int l = blockIdx.y;
    const RowsLayer& ly = a.ly[l];
    const float* W = params + w_offset(d, l) + (int64_t)member * ly.K * ly.N;
    const float* bv = params + b_offset(d, l) + (int64_t)member * ly.N;
    __nv_bfloat16* pm = pack + ly.pack_off;
    const int npair = ly.Ns >> 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ly.Kp * npair; t += gridDim.x * blockDim.x) {
        const int k = t / npair, n = (t - k * npair) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (k <= ly.K) {
            const float* src = k < ly.K ? W + (int64_t)k * ly.N : bv;
            if (n < ly.N) v0 = src[n];
            if (n + 1 < ly.N) v1 = src[n + 1];
        }
        uint32_t hi, lo;
        split2(v0, v1, hi, lo);
        *reinterpret_cast<uint32_t*>(pm + pack_row_off(k, 0, ly.Kp, ly.Ns) + n) = hi;
        *reinterpret_cast<uint32_t*>(pm + pack_row_off(k, 1, ly.Kp, ly.Ns) + n) = lo;
    }
}

// one k-step group of a chunk: NTW n-tiles of this warp x the block's four M tiles
template <int NTW>
__device__ __forceinline__ void rows_chunk(float (&acc)[kMT][kNT][4], int nk, const __nv_bfloat16* Ah,
                                           const __nv_bfloat16* Al, const __nv_bfloat16* Wh, const __nv_bfloat16* Wl,
                                           int a_off, int a_mt_stride, int b_off, int Ns, int warp) {
    for (int kk = 0; kk < nk; ++kk) {
        uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int off = kk * 16 * Ns + b_off + (warp + kWarps * i) * 8;
            ldsm_x2_t(bh[i], smem_u32(Wh + off));
            ldsm_x2_t(bl[i], smem_u32(Wl + off));
        }
#pragma unroll
        for (int mt = 0; mt < kMT; ++mt) {
            uint32_t ah[4], al[4];
            ldsm_x4(ah, smem_u32(Ah + a_off + mt * a_mt_stride + kk * 16));
            ldsm_x4(al, smem_u32(Al + a_off + mt * a_mt_stride + kk * 16));
#pragma unroll
            for (int i = 0; i < NTW; ++i) {
                mma16816(acc[mt][i], al, bh[i]);
                mma16816(acc[mt][i], ah, bl[i]);
                mma16816(acc[mt][i], ah, bh[i]);
            }
        }
    }
}

__global__ void __launch_bounds__(kT, 1) k_mlp_rows(const __grid_constant__ RowsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = a.L;
    // [stage 0 .. kStages-1 | act 0 (hi | lo) | act 1 (hi | lo) | raw input rows x 2]
    auto stage_ptr = [&](int s) { return reinterpret_cast<__nv_bfloat16*>(smem_raw + s * a.stage_bytes); };
    auto act_ptr = [&](int s) { return reinterpret_cast<__nv_bfloat16*>(smem_raw + kStages * a.stage_bytes + s * a.act_bytes); };
    auto raw_ptr = [&](int s) { return reinterpret_cast<float*>(smem_raw + kStages * a.stage_bytes + 2 * a.act_bytes + s * a.raw_bytes); };
    __shared__ __align__(8) uint64_t s_bar[2 * kStages];   // full[], empty[]
    const uint32_t full0 = smem_u32(&s_bar[0]), empty0 = smem_u32(&s_bar[kStages]);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();                                        // the pack kernel of this call
    uint32_t use = 0, use_p = 0;                      // chunks consumed (every thread) / issued (issuer warp)
    auto producer_acquire = [&]() -> int {
        const int s = use_p % kStages;
        if (use_p >= kStages) mbar_wait(empty0 + 8 * s, ((use_p / kStages) & 1) ^ 1);
        ++use_p;
        return s;
    };
    auto consumer_wait = [&]() -> int {
        const int s = use % kStages;
        mbar_wait(full0 + 8 * s, (use / kStages) & 1);
        ++use;
        return s;
    };
    auto consumer_release = [&](int s) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    };

    const int64_t nblk = (a.B + kBlockRows - 1) / kBlockRows;
    const int64_t my_blocks = nblk > (int64_t)blockIdx.x ? (nblk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // chunk stream of this CTA: for every block, for every layer, for every chunk -- issued one chunk ahead
    int64_t ib = 0;
    int il = 0, ic = 0;
    auto issue_w = [&]() {
        if (warp != kIssuer || ib >= my_blocks) return;
        const RowsLayer& ly = a.ly[il];
        const int st = producer_acquire();
        if (lane == 0) {
            const int rows = min(kWChunk, ly.Kp - ic * kWChunk);
            const uint32_t bytes = (uint32_t)rows * ly.Ns * 4;
            const uint32_t fb = full0 + 8 * st;
            mbar_expect_tx(fb, bytes);
            bulk_g2s(smem_u32(stage_ptr(st)), a.pack + ly.pack_off + (int64_t)ic * (kWChunk * ly.Ns * 2), bytes, fb);
        }
        if (++ic >= chunk_count(ly.Kp)) {
            ic = 0;
            if (++il >= L) { il = 0; ++ib; }
        }
    };
    for (int i = 0; i < kStages - 1; ++i) issue_w();
    // raw input rows of a block -> shared memory with cp.async, issued a whole block ahead: the HBM latency of the
    // next block's rows hides behind this block's layers (the loads used to sit, exposed, in front of every block)
    const int K0 = a.ly[0].K;
    auto fetch_rows = [&](int64_t blk, int buf) {
        if (blk >= nblk) return;
        const int64_t row0 = blk * kBlockRows;
        const int64_t n = min((int64_t)kBlockRows, a.B - row0) * K0;
        const float* src = a.x + row0 * K0;
        const uint32_t dst = smem_u32(raw_ptr(buf));
        for (int64_t t = tid; t < n; t += kT)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + (uint32_t)t * 4u), "l"(src + t));
    };
    fetch_rows(blockIdx.x, 0);
    asm volatile("cp.async.commit_group;\n" ::);
    int rbuf = 0;

    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t row0 = blk * kBlockRows;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                                  // this block's rows are in raw[rbuf]; raw[rbuf ^ 1] is free
        fetch_rows(blk + gridDim.x, rbuf ^ 1);
        asm volatile("cp.async.commit_group;\n" ::);
        // ---- input rows -> act 0 as bf16 hi/lo: [x | 1 | 0...] over Kp columns
        {
            const RowsLayer& l0 = a.ly[0];
            __nv_bfloat16* ah = act_ptr(0);
            __nv_bfloat16* al = ah + kBlockRows * a.as_max;
            const int half = l0.Kp / 2;
            for (int t = tid; t < kBlockRows * half; t += kT) {
                const int i = t / half, j = (t - i * half) * 2;
                const int64_t b = row0 + i;
                float v[2] = {0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = j + q;
                    if (c < l0.K) v[q] = b < a.B ? raw_ptr(rbuf)[i * l0.K + c] : 0.f;
                    else if (c == l0.K) v[q] = 1.f;
                }
                uint32_t hi, lo;
                split2(v[0], v[1], hi, lo);
                *reinterpret_cast<uint32_t*>(ah + i * l0.As + j) = hi;
                *reinterpret_cast<uint32_t*>(al + i * l0.As + j) = lo;
            }
        }
        rbuf ^= 1;
        __syncthreads();
        int cur = 0;
        for (int l = 0; l < L; ++l) {
            const RowsLayer& ly = a.ly[l];
            const bool last = l == L - 1;
            const int Ns = ly.Ns, As = ly.As;
            const __nv_bfloat16* Ah = act_ptr(cur);
            const __nv_bfloat16* Al = Ah + kBlockRows * a.as_max;
            float acc[kMT][kNT][4];
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
                for (int i = 0; i < kNT; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mt][i][q] = 0.f;
            const int ntw = ly.nt > warp ? (ly.nt - warp + kWarps - 1) / kWarps : 0;
            const int a_lane_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * As + 8 * (lane >> 4);
            const int b_lane_off = (lane & 15) * Ns;
            const int nchunks = chunk_count(ly.Kp);
            for (int c = 0; c < nchunks; ++c) {
                issue_w();
                const int st = consumer_wait();
                const int rows = min(kWChunk, ly.Kp - c * kWChunk);
                const __nv_bfloat16* Wh = stage_ptr(st);
                const __nv_bfloat16* Wl = Wh + rows * Ns;
                const int ao = a_lane_off + c * kWChunk;
                switch (ntw) {                               // warp-uniform
                case 4: rows_chunk<4>(acc, rows / 16, Ah, Al, Wh, Wl, ao, 16 * As, b_lane_off, Ns, warp); break;
                case 3: rows_chunk<3>(acc, rows / 16, Ah, Al, Wh, Wl, ao, 16 * As, b_lane_off, Ns, warp); break;
                case 2: rows_chunk<2>(acc, rows / 16, Ah, Al, Wh, Wl, ao, 16 * As, b_lane_off, Ns, warp); break;
                case 1: rows_chunk<1>(acc, rows / 16, Ah, Al, Wh, Wl, ao, 16 * As, b_lane_off, Ns, warp); break;
                default: break;
                }
                consumer_release(st);
            }
            // ---- epilogue: activation + hi/lo split into the other buffer (next layer's operand), or the output rows
            __nv_bfloat16* Dh = act_ptr(cur ^ 1);
            __nv_bfloat16* Dl = Dh + kBlockRows * a.as_max;
            const int dstride = last ? 0 : a.ly[l + 1].As;
            const int N = ly.N;
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt) {
#pragma unroll
                for (int i = 0; i < kNT; ++i) {
                    const int tile = warp + kWarps * i;
                    if (tile >= ly.nt) continue;
                    const int n = tile * 8 + 2 * (lane & 3);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = mt * 16 + (lane >> 2) + 8 * h;
                        float v0 = acc[mt][i][2 * h], v1 = acc[mt][i][2 * h + 1];
                        if (last) {
                            const int64_t b = row0 + r;
                            if (b < a.B) {
                                if (n < N) a.out[b * N + n] = v0;
                                if (n + 1 < N) a.out[b * N + n + 1] = v1;
                            }
                        } else {
                            // columns beyond N inside the last tile: column N is the bias column (1), the rest 0
                            v0 = n < N ? apply_act(v0, a.act) : (n == N ? 1.f : 0.f);
                            v1 = n + 1 < N ? apply_act(v1, a.act) : (n + 1 == N ? 1.f : 0.f);
                            uint32_t hi, lo;
                            split2(v0, v1, hi, lo);
                            *reinterpret_cast<uint32_t*>(Dh + r * dstride + n) = hi;
                            *reinterpret_cast<uint32_t*>(Dl + r * dstride + n) = lo;
                        }
                    }
                }
            }
            if (!last) {
                // padding columns of the next operand that no tile wrote: [roundup8(N), Kp_next), column N = 1
                const RowsLayer& nx = a.ly[l + 1];
                const int c0 = (N + 7) & ~7, wpad = nx.Kp - c0;
                for (int t = tid; t < kBlockRows * wpad; t += kT) {
                    const int r = t % kBlockRows, c = c0 + t / kBlockRows;
                    Dh[r * dstride + c] = __float2bfloat16(c == N ? 1.f : 0.f);
                    Dl[r * dstride + c] = __float2bfloat16(0.f);
                }
            }
            __syncthreads();
            cur ^= 1;
        }
    }
}

int rows_plan(const mb_mlp_desc& d, RowsArgs& a, size_t* pack_bytes, size_t* smem) {
    const int L = d.num_layers;
    a.L = L;
    a.act = d.activation;
    int64_t pack_elems = 0;
    int ns_max = 0, as_max = 0;
    for (int l = 0; l < L; ++l) {
        RowsLayer& ly = a.ly[l];
        ly.K = d.sizes[l];
        ly.N = d.sizes[l + 1];
        if (ly.N > 8 * kWarps * kNT || ly.K > 8 * kWarps * kNT) return MB_EUNSUPPORTED;
        ly.Kp = (ly.K + 1 + 15) / 16 * 16;
        ly.Ns = stride_rule(ly.N);
        ly.As = stride_rule(ly.K + 1);
        ly.nt = (ly.N + 7) / 8;
        ly.pack_off = pack_elems;
        pack_elems += (int64_t)2 * ly.Kp * ly.Ns;
        ns_max = ly.Ns > ns_max ? ly.Ns : ns_max;
        as_max = ly.As > as_max ? ly.As : as_max;
    }
    a.as_max = as_max;
    a.stage_bytes = kWChunk * ns_max * 2 * 2;
    a.act_bytes = 2 * kBlockRows * as_max * 2;
    if (pack_bytes) *pack_bytes = ((size_t)pack_elems * 2 + 255) & ~size_t(255);
    a.raw_bytes = (kBlockRows * d.sizes[0] * 4 + 15) & ~15;
    if (smem) *smem = kStages * (size_t)a.stage_bytes + 2 * (size_t)a.act_bytes + 2 * (size_t)a.raw_bytes;
    return MB_OK;
}

}  // namespace

bool mlp_rows_supported(const mb_mlp_desc& d) {
    RowsArgs a{};
    size_t pb = 0, sm = 0;
    return rows_plan(d, a, &pb, &sm) == MB_OK && sm <= 220 * 1024;
}

size_t mlp_rows_pack_bytes(const mb_mlp_desc& d) {
    RowsArgs a{};
    size_t pb = 0;
    return rows_plan(d, a, &pb, nullptr) == MB_OK ? pb : 0;
}

// out[B][N_L] = MLP_member(x[B][K0]) (no activation on the last layer); pack: mlp_rows_pack_bytes of scratch
int launch_mlp_rows_forward(const mb_mlp_desc& d, const float* params, int member, const float* x, int64_t B,
                            float* out, void* pack, cudaStream_t s) {
    RowsArgs a{};
    size_t pb = 0, smem = 0;
    int rc = rows_plan(d, a, &pb, &smem);
    if (rc) { set_error("mlp_rows: widths above 256 are not covered"); return rc; }
    a.x = x;
    a.B = B;
    a.out = out;
    a.pack = reinterpret_cast<const __nv_bfloat16*>(pack);
    launch_pdl(k_rows_pack, dim3(32, (unsigned)a.L), dim3(256), 0, s, d, a, params, member,
               reinterpret_cast<__nv_bfloat16*>(pack));
    MB_CUDA_LAUNCH_CHECK("k_rows_pack");
    static size_t configured[64] = {0};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64 && configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_mlp_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("k_mlp_rows smem %zu: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
        configured[dev] = smem;
    }
    const int64_t nblk = (B + kBlockRows - 1) / kBlockRows;
    const unsigned grid = (unsigned)(nblk < sms ? nblk : sms);
    launch_pdl(k_mlp_rows, dim3(grid), dim3(kT), smem, s, a);
    MB_CUDA_LAUNCH_CHECK("k_mlp_rows");
    return MB_OK;
}

}  // namespace mb
