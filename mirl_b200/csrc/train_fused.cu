// Persistent fused training step of an ensemble MLP: forward + loss head + backward + Adam for
// `nsteps` consecutive steps in ONE cooperative launch.
//
// Replaces, per step, PEModelTrainer.train_from_torch_batch (mirl/agents/mbpo/pe_model_trainer.py:85-106:
// PEModel.compute_loss -> pe_model.py:71-143, loss.backward(), torch.optim.Adam.step over 84 tensors;
// 4 538 ATen calls) and, with the MSE head, the critic half of SACAgent.train_from_torch_batch
// (mirl/agents/sac_agent.py:152-169 -> ddpg_agent.py:138-149, qf_optimizer.step, soft target update
// torch_modules/utils.py:151-155).  Round 1 ran this as ten dependent launches (0.129 ms at 7 x 256).
//
// Shape of the problem: 7 members x 256 rows x 4x200 is 1.4 GFLOP and 26 MB of Adam traffic per step --
// microseconds of roofline time -- behind a chain of 9 dependent GEMMs, so the design minimises
// dependent latency, not tensor-pipe occupancy:
//
//   * grid = members x G CTAs (G = 16), all co-resident (cooperative launch), looping over the steps.
//   * phase A (row split, no communication): CTA (e, r) owns rows [16r, 16r+16) of member e's batch and
//     runs the whole forward, the loss head and the whole dX chain on them with the activations and
//     pre-activations in shared memory.  The per-layer GEMMs are warp-level tensor-core MMAs
//     (mma.sync m16n8k16, bf16 hi/lo split operands, three MMAs per product = fp32-class accuracy,
//     fp32 accumulators): a 16-row block is exactly one MMA M tile.  tcgen05 is the wrong tool at this
//     size: its cost per instruction is max(M,128)*N/256 cycles whatever M is, so 256 rows can occupy at
//     most two SMs per member at full efficiency, and the chain of 9 layers x ~3 us is what round 1
//     measured.  Weights stream from L2 (the bf16 hi/lo "pack", written by phase B) through a
//     double-buffered cp.async ring of 64-row chunks shared by the 8 warps.
//   * member barrier (G CTAs, one global counter, release/acquire).
//   * phase B (output split): the weight-gradient GEMMs dW_l = h_{l-1}^T g_l (reduction over the batch)
//     are cut into strips of 16 weight rows; a CTA owns a few strips of one layer, streams g_l / h_{l-1}
//     (bf16 hi/lo, written by phase A) in 64-row chunks, and applies weight decay + Adam in the
//     accumulator registers: parameters, both moments and the bf16 pack are updated in place; no gradient
//     is ever stored.  The bias rides in the GEMMs as weight row K (activations carry a constant 1
//     column), so db is row K of the same product.
//   * member barrier, next step.  Bootstrap batches are gathered inside the kernel from the device index
//     table (pe_model_trainer.py:179,268), so sufficiently_train issues one launch per report_freq block.
#include <cuda_bf16.h>
#include "common.cuh"
#include "pe_kernels.h"
#include "train_fused.h"
#include "tc_ptx.cuh"

// -DMB_FUSED_TIMELINE: CTA 0 records clock64 stamps of step 2 into the debug page of the workspace
// (profiles/dbg_train_fused.py timeline)
#ifdef MB_FUSED_TIMELINE
#define TL_ON (blockIdx.x == 0 && threadIdx.x == 0 && step == 2)
#define TL_SET(i) do { if (TL_ON) a.dbg[i] = clock64(); } while (0)
#else
#define TL_SET(i)
#endif

namespace mb {

namespace {

constexpr int kT = 256;          // threads per CTA
constexpr int kWarps = 8;
constexpr int kRows = 16;        // batch rows per block (one MMA M tile)
constexpr int kChunk = 64;       // weight rows / batch rows per streamed chunk
constexpr int kMaxNT = 4;        // n-tiles (8 columns) per warp: widths up to 256
constexpr int kMaxStrips = 4;    // 16-row weight strips per phase-B job

// ---- PTX helpers -----------------------------------------------------------------------------------
// ldmatrix reads shared memory: volatile + memory clobber keeps it behind the barriers that publish
// the data.  The MMA is register-only: plain asm, so the compiler interleaves independent chains.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
// D += A(16x16, row) * B(16x8, col), bf16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi): hi*hi + lo*hi + hi*lo carries ~16 mantissa bits
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma16816(c, al, bh);
    mma16816(c, ah, bl);
    mma16816(c, ah, bh);
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
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// G CTAs of one member: arrive (after the CTA's global writes) / wait until `target` arrivals
__device__ __forceinline__ void member_arrive(unsigned* cnt) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1u);
    }
}
__device__ __forceinline__ void member_wait(const unsigned* cnt, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire(cnt) < target) {
            if (++spins > (1u << 26)) __trap();          // never hang the box: a lost CTA is a bug
        }
        // the bulk copies this thread issues next read, through the async proxy, global data that
        // other CTAs wrote with ordinary stores
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
// activation and its derivative from one sigmoid: swish h = z s, h' = s (1 + z (1 - s)) = s + h (1 - s)
__device__ __forceinline__ float act_both(float z, int act, float& deriv) {
    if (act == MB_ACT_RELU) { deriv = z > 0.f ? 1.f : 0.f; return fmaxf(z, 0.f); }
    const float sg = fast_sigmoid(z);
    const float h = z * sg;
    deriv = fmaf(h, 1.f - sg, sg);
    return h;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// bf16 offset of pack element (row k, column 0) of one member's layer pack; `half` 0 = hi, 1 = lo.
// Layout: 64-row chunks, each [hi rows x Ns][lo rows x Ns] contiguous -> a chunk is ONE bulk copy.
__device__ __forceinline__ int pack_row_off(int k, int half, int Kp, int Ns) {
    const int c = k >> 6;
    const int rows = min(kChunk, Kp - c * kChunk);
    return c * (kChunk * Ns * 2) + half * rows * Ns + (k & 63) * Ns;
}

struct StepScalars { float step_size, bc2_sqrt; };

}  // namespace

// ------------------------------------------------------------------------------------------------------
// KS = compile-time bound of the dX reduction length in k-steps (its operand fragments live in registers)
template <int KS>
__global__ void __launch_bounds__(kT, 1)
k_train_fused(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = a.G;
    const int slot = blockIdx.x % G;                  // CTA index inside the member's group
    const int e = a.e0 + blockIdx.x / G;              // ensemble member
    const int L = a.L;
    const int act_kind = a.act;

    // [stage 0 | stage 1 | act 0 | act 1 | z | head | red]
    __nv_bfloat16* const stage0 = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* const stage1 = reinterpret_cast<__nv_bfloat16*>(smem_raw + a.stage_bytes);
    __nv_bfloat16* const act0 = reinterpret_cast<__nv_bfloat16*>(smem_raw + 2 * a.stage_bytes);
    __nv_bfloat16* const act1 = reinterpret_cast<__nv_bfloat16*>(smem_raw + 2 * a.stage_bytes + a.act_bytes);
    float* const zbuf = reinterpret_cast<float*>(smem_raw + 2 * a.stage_bytes + 2 * a.act_bytes);
    float* const headbuf = reinterpret_cast<float*>(smem_raw + 2 * a.stage_bytes + 2 * a.act_bytes + a.z_bytes);
    float* const redbuf = reinterpret_cast<float*>(smem_raw + 2 * a.stage_bytes + 2 * a.act_bytes + a.z_bytes + a.head_bytes);
    auto stage_ptr = [&](int s) { return s ? stage1 : stage0; };
    auto act_ptr = [&](int s) { return s ? act1 : act0; };
    __shared__ StepScalars s_sc;
    __shared__ float s_part[kWarps];
    __shared__ __align__(8) uint64_t s_full[2];       // "chunk landed" barriers of the two stages
    const uint32_t full0 = smem_u32(&s_full[0]);
    uint32_t par0 = 0, par1 = 0;                      // phase parity of each stage's barrier
    if (tid == 0) {
        mbar_init(full0, 1);
        mbar_init(full0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait_stage = [&](int s) {
        if (s) { mbar_wait(full0 + 8, par1); par1 ^= 1; } else { mbar_wait(full0, par0); par0 ^= 1; }
    };

    __nv_bfloat16* const pack = a.pack;
    __nv_bfloat16* const hbuf = a.hbuf;
    __nv_bfloat16* const gbuf = a.gbuf;
    unsigned* const bar = a.barrier + e * 32;         // one 128-byte line per member
    unsigned bar_target = 0;
    const int njobs = a.njobs[slot];

    // ---- prologue: build this CTA's part of the bf16 pack from the fp32 master weights (every column
    // of every row of its strips, padding included, so the pack needs no memset) ----------------------
    for (int j = 0; j < njobs; ++j) {
        const FusedJob jb = a.jobs[slot][j];
        const FusedLayer& ly = a.ly[jb.layer];
        const float* W = a.params + ly.w_off + (int64_t)e * ly.K * ly.N;
        const float* bv = a.params + ly.b_off + (int64_t)e * ly.N;
        __nv_bfloat16* pm = pack + ly.pack_off + (int64_t)e * 2 * ly.Kp * ly.Ns;
        const int k0 = jb.strip0 * 16, k1 = min(k0 + jb.nstrips * 16, ly.Kp);
        const int npair = ly.Ns >> 1;
        for (int t = tid; t < (k1 - k0) * npair; t += kT) {
            const int k = k0 + t / npair, n = (t % npair) * 2;
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
    bar_target += G;
    member_arrive(bar);
    member_wait(bar, bar_target);

    // Adam bias corrections: beta^t carried as running products in double by thread 0
    double b1t = a.beta1_pow0, b2t = a.beta2_pow0;

    for (int step = 0; step < a.nsteps; ++step) {
        // batch of this step inside the epoch schedule (pe_model_trainer.py:240,268)
        const int64_t tglob = a.first_step + step;
        const int64_t kb = tglob % a.steps_per_epoch;
        const int64_t j0 = kb * a.batch_size;
        const int B = (int)min((int64_t)a.batch_size, a.train_length - j0);
        const int nblk = (B + kRows - 1) / kRows;
        const float invB = 1.f / (float)B;
        float* const eloss = a.ensemble_loss + (int64_t)step * a.E;
        float* const lterms = a.loss_terms + (int64_t)step * 3;

        // ================================ phase A ================================================
        TL_SET(0);
        for (int blk = slot; blk < nblk; blk += G) {
            const int row0 = blk * kRows;
            // ---- chunk stream over all layer passes of this block: forward 0..L-1, then dX L-1..1.
            // One elected thread issues ONE bulk copy per chunk (the pack keeps a chunk's hi and lo
            // halves adjacent); the copy of chunk i+1 is in flight while chunk i is consumed.
            int is_pass = 0, is_chunk = 0;            // next chunk to ISSUE (thread 0 only uses these)
            const int npass = 2 * L - 1;
            auto pass_layer = [&](int p) { return p < L ? p : 2 * L - 1 - p; };
            auto issue_next = [&](int stage) {
                if (tid != 0 || is_pass >= npass) return;
                const FusedLayer& ly = a.ly[pass_layer(is_pass)];
                const int r0 = is_chunk * kChunk;
                const int rows = min(kChunk, ly.Kp - r0);
                const __nv_bfloat16* src = pack + ly.pack_off + (int64_t)e * 2 * ly.Kp * ly.Ns + (int64_t)is_chunk * (kChunk * ly.Ns * 2);
#ifdef MB_EXP_NOCOPY
                const uint32_t bytes = 16;
#else
                const uint32_t bytes = (uint32_t)rows * ly.Ns * 4;
#endif
                const uint32_t fb = full0 + 8 * stage;
                mbar_expect_tx(fb, bytes);
                // (one copy per chunk: splitting it into four concurrent pieces measured 20 % slower)
                bulk_g2s(smem_u32(stage_ptr(stage)), src, bytes, fb);
                if (++is_chunk * kChunk >= ly.Kp) { is_chunk = 0; ++is_pass; }
            };
            issue_next(0);

            // ---- stage the normalised inputs [obs | act | 1 | 0...] as bf16 hi/lo (layer 0 operand)
            {
                const FusedLayer& l0 = a.ly[0];
                const int As = l0.As, IN = l0.K, O = a.O;
                __nv_bfloat16* ah = act0;
                __nv_bfloat16* al = ah + kRows * a.as_max;
                for (int t = tid; t < kRows * (As / 2); t += kT) {
                    const int i = t / (As / 2), j = (t % (As / 2)) * 2;
                    const int b = row0 + i;
                    float v[2] = {0.f, 0.f};
                    if (b < B) {
                        const int64_t r = a.index ? a.index[(int64_t)e * a.index_stride + j0 + b]
                                                  : (a.per_member_rows ? (int64_t)e * a.batch_size + b : (int64_t)b);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int c = j + q;
                            if (c < O) {
                                const float x = a.obs[r * O + c];
                                v[q] = a.norm ? (x - a.norm[c]) / (a.norm[O + c] + kNormEps) : x;
                            } else if (c < IN) {
                                const int ca = c - O;
                                const float x = a.actn[r * a.A + ca];
                                v[q] = a.norm ? (x - a.norm[2 * O + ca]) / (a.norm[2 * O + a.A + ca] + kNormEps) : x;
                            } else if (c == IN) {
                                v[q] = 1.f;
                            }
                        }
                    }
                    uint32_t hi, lo;
                    split2(v[0], v[1], hi, lo);
                    *reinterpret_cast<uint32_t*>(ah + i * As + j) = hi;
                    *reinterpret_cast<uint32_t*>(al + i * As + j) = lo;
                }
            }
            __syncthreads();

            int cur = 0;                                  // act[cur] = input operand of the running pass
            int cs_stage = 0;                             // stage holding the chunk to CONSUME
            for (int pass = 0; pass < npass; ++pass) {
                TL_SET(8 + pass);
                const bool fwd = pass < L;
                const int l = pass_layer(pass);
                const FusedLayer& ly = a.ly[l];
                const int Ns = ly.Ns;
                // operand A: forward -> activations (stride ly.As); dX -> g_l (stride Ns)
                const int astride = fwd ? ly.As : Ns;
                const __nv_bfloat16* Ah = act_ptr(cur);
                const __nv_bfloat16* Al = Ah + kRows * a.as_max;
                // copy this pass's A operand to global for phase B (h_l for forward, g_l for dX):
                // coalesced 16-byte pieces of the 16 rows
                {
                    __nv_bfloat16* gh = (fwd ? hbuf + ly.h_off : gbuf + ly.g_off) + ((int64_t)e * a.Bp + row0) * astride;
                    __nv_bfloat16* gl = gh + (int64_t)a.E * a.Bp * astride;
                    const int pieces = kRows * astride / 8;
                    for (int p = tid; p < pieces; p += kT) {
                        *reinterpret_cast<uint4*>(gh + p * 8) = *reinterpret_cast<const uint4*>(Ah + p * 8);
                        *reinterpret_cast<uint4*>(gl + p * 8) = *reinterpret_cast<const uint4*>(Al + p * 8);
                    }
                }
                TL_SET(60 + pass * 6 + 0);
                const int nchunks = (ly.Kp + kChunk - 1) / kChunk;
                float acc[kMaxNT][4];
#pragma unroll
                for (int i = 0; i < kMaxNT; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
                const int nt_total = fwd ? ly.nt_out : ly.nt_in;   // output n-tiles of this pass
                const int out_w = fwd ? ly.N : ly.K;               // real output columns
                // destination operand buffer of this pass
                __nv_bfloat16* Dh = act_ptr(cur ^ 1);
                __nv_bfloat16* Dl = Dh + kRows * a.as_max;
                const bool last_fwd = fwd && l == L - 1;
                // stride of the produced operand: forward -> next layer's input stride; dX -> Ns of layer l-1
                const int dstride = fwd ? (last_fwd ? 0 : a.ly[l + 1].As) : a.ly[l - 1].Ns;
                const float* zin = zbuf + (fwd ? l : l - 1) * kRows * a.z_stride;

                auto epilogue_tile = [&](int tile, const float (&c)[4]) {
                    // accumulator fragment: rows lane/4 and lane/4+8, columns n0 + 2*(lane%4), +1
                    const int n = tile * 8 + 2 * (lane & 3);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int i = (lane >> 2) + 8 * h;
                        float v0 = c[2 * h], v1 = c[2 * h + 1];
                        if (fwd) {
                            if (last_fwd) {
                                if (n < out_w) headbuf[i * a.head_stride + n] = v0;
                                if (n + 1 < out_w) headbuf[i * a.head_stride + n + 1] = v1;
                                continue;
                            }
                            // the backward pass needs act'(z), not z: keep that (one sigmoid serves both)
                            float* zr = const_cast<float*>(zin) + i * a.z_stride + n;
                            float d0 = 0.f, d1 = 0.f;
                            if (n < out_w) v0 = act_both(v0, act_kind, d0); else v0 = (n == out_w) ? 1.f : 0.f;
                            if (n + 1 < out_w) v1 = act_both(v1, act_kind, d1); else v1 = (n + 1 == out_w) ? 1.f : 0.f;
                            *reinterpret_cast<float2*>(zr) = make_float2(d0, d1);
                        } else {
                            const float2 d = *reinterpret_cast<const float2*>(zin + i * a.z_stride + n);
                            v0 = n < out_w ? v0 * d.x : 0.f;
                            v1 = n + 1 < out_w ? v1 * d.y : 0.f;
                        }
                        uint32_t hi, lo;
                        split2(v0, v1, hi, lo);
                        *reinterpret_cast<uint32_t*>(Dh + i * dstride + n) = hi;
                        *reinterpret_cast<uint32_t*>(Dl + i * dstride + n) = lo;
                    }
                };

                // zero the destination's padding columns beyond the tiles the epilogue writes
                if (!last_fwd) {
                    const int c0 = nt_total * 8;
                    const int wpad = dstride - c0;
                    for (int t = tid; t < kRows * wpad; t += kT) {
                        const int i = t & (kRows - 1), c = c0 + (t >> 4);
                        // forward: column N may fall beyond the last tile (N multiple of 8): it is the bias 1
                        const float v = (fwd && c == out_w) ? 1.f : 0.f;
                        Dh[i * dstride + c] = __float2bfloat16(v);
                        Dl[i * dstride + c] = __float2bfloat16(0.f);
                    }
                }

                TL_SET(60 + pass * 6 + 1);
                // per-lane ldmatrix offsets (elements) that do not depend on the chunk
                const int a_lane_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * astride + 8 * (lane >> 4);
                // dX: every chunk needs the whole g_l operand (the reduction runs over its columns): the
                // fragments of all k-steps are loaded ONCE per pass and stay in registers -- re-reading
                // them per chunk made the pass shared-memory-bandwidth bound (8 warps x 4 chunks x 13 KB)
                uint32_t gfh[KS][4], gfl[KS][4];
                if (!fwd) {
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        if (kk < ly.ksteps_out) {
                            ldsm_x4(gfh[kk], smem_u32(Ah + a_lane_off + kk * 16));
                            ldsm_x4(gfl[kk], smem_u32(Al + a_lane_off + kk * 16));
                        }
                    }
                }
                for (int c = 0; c < nchunks; ++c) {
                    issue_next(cs_stage ^ 1);
                    wait_stage(cs_stage);
                    if (c == 0) TL_SET(60 + pass * 6 + 2);
                    const int rows = min(kChunk, ly.Kp - c * kChunk);
                    const __nv_bfloat16* Wh = stage_ptr(cs_stage);
                    const __nv_bfloat16* Wl = Wh + rows * Ns;
#ifdef MB_EXP_NOMMA
                    if (false) {} else if (a.nsteps < 0)
#endif
                    if (fwd) {
                        // reduction over the chunk's rows; warp owns n-tiles warp, warp+8, ...
                        const int b_lane_off = (lane & 15) * Ns;
                        const int nk = rows / 16;
                        // fully unrolled (guarded) so that the k-steps' fragment loads are scheduled ahead of
                        // the previous k-step's MMAs
#pragma unroll
                        for (int kk = 0; kk < kChunk / 16; ++kk) {
                            if (kk >= nk) break;
                            const int kcol = c * kChunk + kk * 16;
                            uint32_t ah[4], al[4];
                            ldsm_x4(ah, smem_u32(Ah + a_lane_off + kcol));
                            ldsm_x4(al, smem_u32(Al + a_lane_off + kcol));
                            uint32_t bh[kMaxNT][2], bl[kMaxNT][2];
#pragma unroll
                            for (int i = 0; i < kMaxNT; ++i) {
                                const int tile = warp + kWarps * i;
                                if (tile < nt_total) {
                                    const int off = kk * 16 * Ns + b_lane_off + tile * 8;
                                    ldsm_x2_t(bh[i], smem_u32(Wh + off));
                                    ldsm_x2_t(bl[i], smem_u32(Wl + off));
                                }
                            }
#pragma unroll
                            for (int i = 0; i < kMaxNT; ++i) {
                                const int tile = warp + kWarps * i;
                                if (tile < nt_total) mma3(acc[i], ah, al, bh[i], bl[i]);
                            }
                        }
                    } else {
                        // chunk rows = output columns k; warp owns output tile 8c + warp; full reduction over n.
                        // six independent accumulator chains (3 terms x even/odd k-step): one tile per warp
                        // would otherwise be a 39-deep dependent MMA chain
                        const int tile = c * (kChunk / 8) + warp;
                        if (tile < nt_total) {
                            float c6[6][4];
#pragma unroll
                            for (int i = 0; i < 6; ++i)
#pragma unroll
                                for (int q = 0; q < 4; ++q) c6[i][q] = 0.f;
                            const int b_lane_off = (warp * 8 + (lane & 7)) * Ns + 8 * ((lane >> 3) & 1);
#pragma unroll
                            for (int kk = 0; kk < KS; ++kk) {
                                if (kk < ly.ksteps_out) {
                                    uint32_t bh[2], bl[2];
                                    ldsm_x2(bh, smem_u32(Wh + b_lane_off + kk * 16));
                                    ldsm_x2(bl, smem_u32(Wl + b_lane_off + kk * 16));
                                    const int o = (kk & 1) * 3;
                                    mma16816(c6[o + 0], gfl[kk], bh);
                                    mma16816(c6[o + 1], gfh[kk], bl);
                                    mma16816(c6[o + 2], gfh[kk], bh);
                                }
                            }
                            float cacc[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                cacc[q] = ((c6[0][q] + c6[3][q]) + (c6[1][q] + c6[4][q])) + (c6[2][q] + c6[5][q]);
                            epilogue_tile(tile, cacc);
                        }
                    }
                    __syncthreads();                       // stage may be refilled by the next issue
                    cs_stage ^= 1;
                }
                TL_SET(60 + pass * 6 + 3);
                if (fwd) {
#pragma unroll
                    for (int i = 0; i < kMaxNT; ++i) {
                        const int tile = warp + kWarps * i;
                        if (tile < nt_total) epilogue_tile(tile, acc[i]);
                    }
                }
                TL_SET(60 + pass * 6 + 4);
                __syncthreads();
                TL_SET(60 + pass * 6 + 5);

                if (last_fwd) {
                    // ---- loss head: writes g_L (bf16 hi/lo, stride Ns) into the destination buffer
                    const int D = a.D, O = a.O;
                    const int gs = ly.Ns;
                    float part = 0.f;
                    // per-(row, column) contributions to the log-std bound gradients: summed in a fixed
                    // order below (and over the row blocks in phase B), so a step is bit-reproducible
                    for (int t = tid; t < 2 * kRows * 64; t += kT) redbuf[t] = 0.f;
                    for (int t = tid; t < kRows * gs; t += kT) {          // clear (pads must be exact zeros)
                        Dh[t] = __float2bfloat16(0.f);
                        Dl[t] = __float2bfloat16(0.f);
                    }
                    __syncthreads();
                    if (a.head_kind == FUSED_HEAD_NLL) {
                        const float* mxp = a.params + a.maxls_off + (int64_t)e * D;
                        const float* mnp = a.params + a.minls_off + (int64_t)e * D;
                        const float* dmean = a.norm + 2 * O + 2 * a.A;
                        const float* dstd = dmean + O;
                        for (int t = tid; t < kRows * D; t += kT) {
                            const int i = t / D, j = t - i * D;
                            const int b = row0 + i;
                            if (b >= B) continue;
                            const int64_t r = a.index ? a.index[(int64_t)e * a.index_stride + j0 + b]
                                                      : (a.per_member_rows ? (int64_t)e * a.batch_size + b : (int64_t)b);
                            const float mu = headbuf[i * a.head_stride + j], raw = headbuf[i * a.head_stride + D + j];
                            const float mx = __ldcg(mxp + j), mn = __ldcg(mnp + j);   // updated by another SM every step
                            const float aa = mx - raw;
                            const float ls1 = mx - softplusf_(aa);
                            const float cc = ls1 - mn;
                            const float ls = mn + softplusf_(cc);
                            float tgt, coef;
                            if (j < O) { tgt = (a.delta[r * O + j] - dmean[j]) / (dstd[j] + kNormEps); coef = 1.f; }
                            else { tgt = a.reward[r]; coef = a.cfg.reward_coef; }
                            const float mask = a.cfg.ignore_terminal_state ? 1.f - a.done[r] : 1.f;
                            const float inv_var = expf(-2.f * ls);
                            const float diff = tgt - mu;
                            const float w = mask * coef * invB;
                            part += w * 0.5f * (kLog2Pi + 2.f * ls + diff * diff * inv_var);
                            const float g_ls = w * (1.f - diff * diff * inv_var);
                            const float sa = softplus_gradf_(aa), sb = softplus_gradf_(cc);
                            const float gm = -w * diff * inv_var, gr = g_ls * sb * sa;
                            const __nv_bfloat16 gmh = __float2bfloat16(gm), grh = __float2bfloat16(gr);
                            Dh[i * gs + j] = gmh;
                            Dl[i * gs + j] = __float2bfloat16(gm - __bfloat162float(gmh));
                            Dh[i * gs + D + j] = grh;
                            Dl[i * gs + D + j] = __float2bfloat16(gr - __bfloat162float(grh));
                            redbuf[i * 64 + j] = g_ls * sb * (1.f - sa);
                            redbuf[(kRows + i) * 64 + j] = g_ls * (1.f - sb);
                        }
                    } else {       // FUSED_HEAD_MSE: F.mse_loss over [E,B,N] against a target shared by the members
                        const int N = ly.N;
                        const float wgt = 1.f / ((float)a.E_total * (float)B * (float)N);
                        for (int t = tid; t < kRows * N; t += kT) {
                            const int i = t / N, j = t - i * N;
                            const int b = row0 + i;
                            if (b >= B) continue;
                            const float q = headbuf[i * a.head_stride + j];
                            const float diff = q - a.target[(int64_t)b * N + j];
                            part += diff * diff * wgt;
                            const float g = 2.f * diff * wgt;
                            const __nv_bfloat16 gh = __float2bfloat16(g);
                            Dh[i * gs + j] = gh;
                            Dl[i * gs + j] = __float2bfloat16(g - __bfloat162float(gh));
                        }
                    }
                    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if (lane == 0) s_part[warp] = part;
                    __syncthreads();
                    float* slot_out = a.block_part + ((int64_t)e * a.nblk_max + blk) * kFusedPartStride;
                    if (tid == 0) {
                        float s = 0.f;
                        for (int w = 0; w < kWarps; ++w) s += s_part[w];
                        slot_out[0] = s;
                    }
                    if (a.head_kind == FUSED_HEAD_NLL && tid < 2 * 64 && (tid & 63) < a.D) {
                        const int which = tid >> 6, j = tid & 63;
                        float s = 0.f;
                        for (int i = 0; i < kRows; ++i) s += redbuf[(which * kRows + i) * 64 + j];
                        slot_out[4 + which * 64 + j] = s;
                    }
                    __syncthreads();
                }
                cur ^= 1;
            }
            // the last dX pass produced g_0 in act[cur]: phase B needs it too
            {
                const FusedLayer& l0 = a.ly[0];
                const int astride = l0.Ns;
                const __nv_bfloat16* Ah = act_ptr(cur);
                const __nv_bfloat16* Al = Ah + kRows * a.as_max;
                __nv_bfloat16* gh = gbuf + l0.g_off + ((int64_t)e * a.Bp + row0) * astride;
                __nv_bfloat16* gl = gh + (int64_t)a.E * a.Bp * astride;
                const int pieces = kRows * astride / 8;
                for (int p = tid; p < pieces; p += kT) {
                    *reinterpret_cast<uint4*>(gh + p * 8) = *reinterpret_cast<const uint4*>(Ah + p * 8);
                    *reinterpret_cast<uint4*>(gl + p * 8) = *reinterpret_cast<const uint4*>(Al + p * 8);
                }
            }
            __syncthreads();
        }
        TL_SET(1);
        bar_target += G;
        member_arrive(bar);
        member_wait(bar, bar_target);
        TL_SET(2);

        // ================================ phase B ================================================
        if (tid == 0) {
            b1t *= (double)a.cfg.beta1;
            b2t *= (double)a.cfg.beta2;
            s_sc.step_size = (float)((double)a.cfg.lr / (1.0 - b1t));
            s_sc.bc2_sqrt = (float)sqrt(1.0 - b2t);
        }
        __syncthreads();
        const float step_size = s_sc.step_size, inv_bc2 = 1.f / s_sc.bc2_sqrt;
        const float beta1 = a.cfg.beta1, beta2 = a.cfg.beta2, eps = a.cfg.eps;
        // torch.optim.Adam, single-tensor rule (pe_model_trainer.py:47-51): returns the new parameter
        auto adam = [&](float p, float& m1, float& m2, float g) -> float {
            m1 = m1 * beta1 + g * (1.f - beta1);
            m2 = m2 * beta2 + (g * g) * (1.f - beta2);
            const float denom = fmaf(fast_sqrt(m2), inv_bc2, eps);
            return fmaf(-step_size, __fdividef(m1, denom), p);
        };

        float l2_part = 0.f;
        for (int j = 0; j < njobs; ++j) {
            const FusedJob jb = a.jobs[slot][j];
            const int l = jb.layer;
            const FusedLayer& ly = a.ly[l];
            const int Ns = ly.Ns, As = ly.As, ns = jb.nstrips;
            const int HT = 16 * ns + 8;                       // h tile stride (odd multiple of 8: conflict-free)
            const __nv_bfloat16* gh_src = gbuf + ly.g_off + (int64_t)e * a.Bp * Ns;
            const __nv_bfloat16* gl_src = gh_src + (int64_t)a.E * a.Bp * Ns;
            const __nv_bfloat16* hh_src = hbuf + ly.h_off + (int64_t)e * a.Bp * As + jb.strip0 * 16;
            const __nv_bfloat16* hl_src = hh_src + (int64_t)a.E * a.Bp * As;
            // g chunk: two bulk copies (hi, lo rows are contiguous); h tile: a few cp.async per thread
            auto issue_b = [&](int c, int stage) {
                const int rows = min(kChunk, nblk * kRows - c * kChunk);
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)rows * Ns * 2;
                    const uint32_t fb = full0 + 8 * stage;
                    const int64_t goff = (int64_t)c * kChunk * Ns;
                    mbar_expect_tx(fb, 2 * bytes);
                    bulk_g2s(smem_u32(stage_ptr(stage)), gh_src + goff, bytes, fb);
                    bulk_g2s(smem_u32(stage_ptr(stage) + kChunk * Ns), gl_src + goff, bytes, fb);
                }
                __nv_bfloat16* th = act_ptr(stage);
                __nv_bfloat16* tl = th + kChunk * HT;
                const int ppr = 2 * ns;                        // 16-byte pieces per row
                for (int p = tid; p < rows * ppr; p += kT) {
                    const int r = p / ppr, q = p - r * ppr;
                    const int64_t so = ((int64_t)c * kChunk + r) * As + q * 8;
                    cp_async16(th + r * HT + q * 8, hh_src + so);
                    cp_async16(tl + r * HT + q * 8, hl_src + so);
                }
                cp_async_commit();
            };
            float acc[kMaxStrips][kMaxNT][4];
#pragma unroll
            for (int s = 0; s < kMaxStrips; ++s)
#pragma unroll
                for (int i = 0; i < kMaxNT; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[s][i][q] = 0.f;
            const int nchunks = (nblk * kRows + kChunk - 1) / kChunk;
            const int nt_total = ly.nt_out;
            issue_b(0, 0);
            for (int c = 0; c < nchunks; ++c) {
                const int stage = c & 1;
                if (c + 1 < nchunks) { issue_b(c + 1, stage ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
                wait_stage(stage);
                __syncthreads();                              // everybody's cp.async pieces of the h tile
                const __nv_bfloat16* Gh = stage_ptr(stage);
                const __nv_bfloat16* Gl = Gh + kChunk * Ns;
                const __nv_bfloat16* Hh = act_ptr(stage);
                const __nv_bfloat16* Hl = Hh + kChunk * HT;
                const int groups = min(kChunk / 16, nblk - c * (kChunk / 16));
                const int g_lane_off = (lane & 15) * Ns;
                const int h_lane_off = ((lane & 7) + 8 * (lane >> 4)) * HT + 8 * ((lane >> 3) & 1);
#pragma unroll
                for (int kk = 0; kk < kChunk / 16; ++kk) {
                    if (kk >= groups) break;
                    uint32_t bh[kMaxNT][2], bl[kMaxNT][2];
#pragma unroll
                    for (int i = 0; i < kMaxNT; ++i) {
                        const int tile = warp + kWarps * i;
                        if (tile < nt_total) {
                            const int off = kk * 16 * Ns + g_lane_off + tile * 8;
                            ldsm_x2_t(bh[i], smem_u32(Gh + off));
                            ldsm_x2_t(bl[i], smem_u32(Gl + off));
                        }
                    }
#pragma unroll
                    for (int s = 0; s < kMaxStrips; ++s) {
                        if (s < ns) {
                            uint32_t ah[4], al[4];
                            const int off = kk * 16 * HT + h_lane_off + s * 16;
                            ldsm_x4_t(ah, smem_u32(Hh + off));
                            ldsm_x4_t(al, smem_u32(Hl + off));
#pragma unroll
                            for (int i = 0; i < kMaxNT; ++i) {
                                const int tile = warp + kWarps * i;
                                if (tile < nt_total) mma3(acc[s][i], ah, al, bh[i], bl[i]);
                            }
                        }
                    }
                }
                __syncthreads();
            }
            TL_SET(40 + j);
            // ---- epilogue: the accumulators go through shared memory (the chunk ring is idle now) so that
            // weight decay + Adam + the pack refresh run as a row-contiguous, 16-byte-vectorised sweep over
            // parameters, moments and pack (fragment-order accesses were 32-byte scattered: the sweep was
            // bound by memory transactions, not bytes)
            float* dW = reinterpret_cast<float*>(stage0);          // [16 * ns][Ns] fp32
#pragma unroll
            for (int s = 0; s < kMaxStrips; ++s) {
                if (s >= ns) continue;
#pragma unroll
                for (int i = 0; i < kMaxNT; ++i) {
                    const int tile = warp + kWarps * i;
                    if (tile >= nt_total) continue;
                    const int n = tile * 8 + 2 * (lane & 3);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = s * 16 + (lane >> 2) + 8 * h;
                        *reinterpret_cast<float2*>(dW + r * Ns + n) = make_float2(acc[s][i][2 * h], acc[s][i][2 * h + 1]);
                    }
                }
            }
            __syncthreads();
            {
                float* Wp = a.params + ly.w_off + (int64_t)e * ly.K * ly.N;
                float* W1 = a.exp_avg + ly.w_off + (int64_t)e * ly.K * ly.N;
                float* W2 = a.exp_avg_sq + ly.w_off + (int64_t)e * ly.K * ly.N;
                float* Bq = a.params + ly.b_off + (int64_t)e * ly.N;
                float* B1 = a.exp_avg + ly.b_off + (int64_t)e * ly.N;
                float* B2 = a.exp_avg_sq + ly.b_off + (int64_t)e * ly.N;
                __nv_bfloat16* pm = pack + ly.pack_off + (int64_t)e * 2 * ly.Kp * Ns;
                const int N = ly.N;
                const int krows = min(16 * ns, ly.K + 1 - jb.strip0 * 16);      // rows k <= K of this job
                const bool vec4 = (N & 3) == 0;
                for (int r = warp; r < krows; r += kWarps) {                    // one weight row per warp
                    const int k = jb.strip0 * 16 + r;
                    const bool isw = k < ly.K;
                    float* rp = isw ? Wp + (int64_t)k * N : Bq;
                    float* r1 = isw ? W1 + (int64_t)k * N : B1;
                    float* r2 = isw ? W2 + (int64_t)k * N : B2;
                    const float wd = isw ? ly.wd : 0.f;
                    __nv_bfloat16* ph = pm + pack_row_off(k, 0, ly.Kp, Ns);
                    __nv_bfloat16* pl = pm + pack_row_off(k, 1, ly.Kp, Ns);
                    const float* dr = dW + r * Ns;
                    if (vec4) {
                        for (int n = lane * 4; n < N; n += 128) {
                            const float4 w = *reinterpret_cast<const float4*>(rp + n);
                            float4 m1 = *reinterpret_cast<const float4*>(r1 + n);
                            float4 m2 = *reinterpret_cast<const float4*>(r2 + n);
                            const float4 g = *reinterpret_cast<const float4*>(dr + n);
                            l2_part += 0.5f * wd * ((w.x * w.x + w.y * w.y) + (w.z * w.z + w.w * w.w));
                            float4 nw;
                            nw.x = adam(w.x, m1.x, m2.x, fmaf(wd, w.x, g.x));
                            nw.y = adam(w.y, m1.y, m2.y, fmaf(wd, w.y, g.y));
                            nw.z = adam(w.z, m1.z, m2.z, fmaf(wd, w.z, g.z));
                            nw.w = adam(w.w, m1.w, m2.w, fmaf(wd, w.w, g.w));
                            *reinterpret_cast<float4*>(rp + n) = nw;
                            *reinterpret_cast<float4*>(r1 + n) = m1;
                            *reinterpret_cast<float4*>(r2 + n) = m2;
                            uint2 hi, lo;
                            split2(nw.x, nw.y, hi.x, lo.x);
                            split2(nw.z, nw.w, hi.y, lo.y);
                            *reinterpret_cast<uint2*>(ph + n) = hi;
                            *reinterpret_cast<uint2*>(pl + n) = lo;
                        }
                    } else {
                        for (int n = lane; n < N; n += 32) {
                            const float w = rp[n];
                            float m1 = r1[n], m2 = r2[n];
                            l2_part += 0.5f * wd * w * w;
                            const float nw = adam(w, m1, m2, fmaf(wd, w, dr[n]));
                            rp[n] = nw;
                            r1[n] = m1;
                            r2[n] = m2;
                            const __nv_bfloat16 h = __float2bfloat16(nw);
                            ph[n] = h;
                            pl[n] = __float2bfloat16(nw - __bfloat162float(h));
                        }
                    }
                }
            }
            __syncthreads();                                  // the next job's first chunk overwrites the ring
        }
        // L2 term of the loss (mlp.py:170-178), from the pre-update weights
        for (int o = 16; o; o >>= 1) l2_part += __shfl_xor_sync(0xffffffffu, l2_part, o);
        if (lane == 0 && l2_part != 0.f) atomicAdd(lterms + 1, l2_part);
        // per-member loss and log-std bounds (pe_model.py:62-69,105-108): the row blocks' partials are
        // summed in block order; Adam on 2 x D scalars per member + the bound term of the loss
        if (slot == G - 1 && warp < 3) {
            const float* parts = a.block_part + (int64_t)e * a.nblk_max * kFusedPartStride;
            if (warp == 0) {
                // lane bk loads block bk's partial: all loads in flight at once, then a fixed-order tree
                float s = 0.f;
                for (int bk = lane; bk < nblk; bk += 32) s += __ldcg(parts + (int64_t)bk * kFusedPartStride);
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) {
                    eloss[e] = s;
                    atomicAdd(lterms + 0, s);
                }
            } else if (a.head_kind == FUSED_HEAD_NLL) {
                const int which = warp - 1;                       // 0: max_logstd, 1: min_logstd
                float bsum = 0.f;
                for (int j = lane; j < a.D; j += 32) {
                    float g4[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int bk = 0; bk < nblk; bk += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (bk + u < nblk) g4[u] += __ldcg(parts + (int64_t)(bk + u) * kFusedPartStride + 4 + which * 64 + j);
                    }
                    float g = (g4[0] + g4[1]) + (g4[2] + g4[3]);
                    g += which ? -a.cfg.bound_reg_coef : a.cfg.bound_reg_coef;
                    const int64_t o = (which ? a.minls_off : a.maxls_off) + (int64_t)e * a.D + j;
                    const float p0 = a.params[o];
                    bsum += (which ? -a.cfg.bound_reg_coef : a.cfg.bound_reg_coef) * p0;
                    a.params[o] = adam(p0, a.exp_avg[o], a.exp_avg_sq[o], g);
                }
                for (int o = 16; o; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
                if (lane == 0) atomicAdd(lterms + 2, bsum);
            }
        }
        // soft target update fused behind Adam (torch_modules/utils.py:151-155): target <- (1-tau) target + tau p
        if (a.target_params != nullptr) {
            __syncthreads();
            for (int j = 0; j < njobs; ++j) {
                const FusedJob jb = a.jobs[slot][j];
                const FusedLayer& ly = a.ly[jb.layer];
                const int k0 = jb.strip0 * 16, k1 = min(k0 + jb.nstrips * 16, ly.K + 1);
                for (int t = tid; t < (k1 - k0) * ly.N; t += kT) {
                    const int k = k0 + t / ly.N, n = t % ly.N;
                    const int64_t o = k < ly.K ? ly.w_off + (int64_t)e * ly.K * ly.N + (int64_t)k * ly.N + n
                                               : ly.b_off + (int64_t)e * ly.N + n;
                    a.target_params[o] = a.target_params[o] * (1.f - a.tau) + a.params[o] * a.tau;
                }
            }
        }
        TL_SET(3);
        bar_target += G;
        member_arrive(bar);
        member_wait(bar, bar_target);
        TL_SET(4);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
static int stride_rule(int n) {           // >= roundup16(n), congruent 8 mod 16: conflict-free ldmatrix rows
    int s = (n + 15) / 16 * 16;
    return s + 8;
}

bool fused_train_supported(const mb_mlp_desc& d, int ne) {
    if (ne < 1 || ne > 148) return false;
    for (int l = 0; l <= d.num_layers; ++l)
        if (d.sizes[l] > 8 * kWarps * kMaxNT) return false;          // widths up to 256
    return true;
}

static int fused_group(int ne, int nsm) {
    int G = nsm / ne;
    if (G > kFusedMaxG) G = kFusedMaxG;
    return G < 1 ? 1 : G;
}

int fused_plan(const mb_mlp_desc& d, int ne, int64_t batch_size, int nsm, FusedArgs& a, FusedSizes& sz) {
    const int L = d.num_layers, E = d.ensemble_size;
    a.L = L;
    a.E = E;
    a.ne = ne;
    a.act = d.activation;
    a.G = fused_group(ne, nsm);
    a.Bp = (int)((batch_size + kChunk - 1) / kChunk * kChunk);
    int64_t pack_elems = 0, h_elems = 0, g_elems = 0;
    int ns_max = 0, as_max = 0, zmax = 0;
    for (int l = 0; l < L; ++l) {
        FusedLayer& ly = a.ly[l];
        ly.K = d.sizes[l];
        ly.N = d.sizes[l + 1];
        ly.Kp = (ly.K + 1 + 15) / 16 * 16;
        ly.Ns = stride_rule(ly.N);
        ly.As = stride_rule(ly.K + 1);
        ly.nt_out = (ly.N + 7) / 8;
        ly.nt_in = (ly.K + 7) / 8;
        ly.ksteps_out = (ly.N + 15) / 16;
        ly.w_off = w_offset(d, l);
        ly.b_off = b_offset(d, l);
        ly.pack_off = pack_elems;
        pack_elems += (int64_t)E * 2 * ly.Kp * ly.Ns;
        ly.h_off = h_elems;
        h_elems += (int64_t)2 * E * a.Bp * ly.As;
        ly.g_off = g_elems;
        g_elems += (int64_t)2 * E * a.Bp * ly.Ns;
        ns_max = ly.Ns > ns_max ? ly.Ns : ns_max;
        as_max = ly.As > as_max ? ly.As : as_max;
        as_max = ly.Ns > as_max ? ly.Ns : as_max;            // act buffers also hold g operands
        if (l + 1 < L) zmax = ly.N > zmax ? ly.N : zmax;
    }
    a.as_max = as_max;
    a.z_stride = (zmax + 7) / 8 * 8 + 4;          // act'(z) rows: float2 stores of whole 8-column tiles
    a.head_stride = d.sizes[L] + 1;
    a.stage_bytes = kChunk * ns_max * 2 * 2;
    // act buffers: phase A [hi|lo] x 16 x as_max; phase B h tiles [hi|lo] x 64 x (16*kMaxStrips+8)
    int act_b = 2 * kRows * as_max * 2;
    const int htile_b = 2 * kChunk * (16 * kMaxStrips + 8) * 2;
    a.act_bytes = act_b > htile_b ? act_b : htile_b;
    a.z_bytes = (L > 1 ? (L - 1) : 1) * kRows * a.z_stride * 4;
    a.head_bytes = (kRows * a.head_stride * 4 + 15) / 16 * 16;
    a.nblk_max = a.Bp / kRows;
    sz.part_bytes = (size_t)E * a.nblk_max * kFusedPartStride * 4;
    sz.smem = 2 * (size_t)a.stage_bytes + 2 * (size_t)a.act_bytes + a.z_bytes + a.head_bytes + 2 * kRows * 64 * 4;
    sz.pack_bytes = (size_t)pack_elems * 2;
    sz.h_bytes = (size_t)h_elems * 2;
    sz.g_bytes = (size_t)g_elems * 2;
    // ---- phase-B jobs: strips of 16 weight rows, a job = up to kMaxStrips strips of one layer
    struct J { int layer, s0, n, cost; };
    J jobs[256];
    int nj = 0;
    // Every job has a fixed cost (pipeline fill, epilogue sweep), so aim at ONE job per CTA: start from the
    // fewest jobs the accumulator budget allows (kMaxStrips strips each) and keep splitting the layer that
    // owns the most expensive job until there are G jobs.
    int pieces[MB_MAX_LAYERS], njtot = 0;
    for (int l = 0; l < L; ++l) {
        pieces[l] = (a.ly[l].Kp / 16 + kMaxStrips - 1) / kMaxStrips;
        njtot += pieces[l];
    }
    while (njtot < a.G) {
        int best = -1, bestc = -1;
        for (int l = 0; l < L; ++l) {
            const int strips = a.ly[l].Kp / 16;
            if (pieces[l] >= strips) continue;
            const int c = (strips + pieces[l] - 1) / pieces[l] * a.ly[l].nt_out;
            if (c > bestc) { bestc = c; best = l; }
        }
        if (best < 0) break;
        ++pieces[best];
        ++njtot;
    }
    for (int l = 0; l < L; ++l) {
        const int strips = a.ly[l].Kp / 16, c1 = a.ly[l].nt_out;
        for (int p = 0; p < pieces[l]; ++p) {                 // as even as possible
            const int s0 = (int)((int64_t)strips * p / pieces[l]), s1 = (int)((int64_t)strips * (p + 1) / pieces[l]);
            if (nj >= 256) return MB_EUNSUPPORTED;
            jobs[nj++] = {l, s0, s1 - s0, 25 + (s1 - s0) * c1};
        }
    }
    // longest-processing-time-first assignment
    for (int i = 0; i < nj; ++i)
        for (int j = i + 1; j < nj; ++j)
            if (jobs[j].cost > jobs[i].cost) { J t = jobs[i]; jobs[i] = jobs[j]; jobs[j] = t; }
    int load[kFusedMaxG] = {0};
    for (int g = 0; g < kFusedMaxG; ++g) a.njobs[g] = 0;
    for (int i = 0; i < nj; ++i) {
        int best = 0;
        for (int g = 1; g < a.G; ++g)
            if (load[g] < load[best]) best = g;
        if (a.njobs[best] >= kFusedMaxJobs) return MB_EUNSUPPORTED;
        FusedJob& fj = a.jobs[best][a.njobs[best]++];
        fj.layer = (short)jobs[i].layer;
        fj.strip0 = (short)jobs[i].s0;
        fj.nstrips = (short)jobs[i].n;
        fj.pad = 0;
        load[best] += jobs[i].cost;
    }
    return MB_OK;
}

size_t fused_train_workspace_bytes(const mb_mlp_desc& d, int64_t batch_size) {
    FusedArgs a{};
    FusedSizes sz{};
    if (fused_plan(d, 1, batch_size, 148, a, sz) != MB_OK) return 0;
    return sz.pack_bytes + sz.h_bytes + sz.g_bytes + 8192 /*barriers + debug page*/ + sz.part_bytes + 1024;
}

int launch_train_fused(const mb_mlp_desc& d, FusedArgs& a, void* ws, size_t ws_bytes, cudaStream_t s) {
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    FusedSizes sz{};
    int rc = fused_plan(d, a.ne, a.batch_size, nsm, a, sz);
    if (rc) { set_error("train_fused: shape not supported"); return rc; }
    const size_t need = sz.pack_bytes + sz.h_bytes + sz.g_bytes + 8192 + sz.part_bytes;
    if (ws == nullptr || ws_bytes < need) {
        set_error("train_fused: workspace %zu < %zu bytes", ws_bytes, need);
        return MB_EWORKSPACE;
    }
    unsigned char* p = reinterpret_cast<unsigned char*>(ws);
    // [barriers 4 KB | debug 4 KB | block partials | pack | h | g]: the first three are zeroed per launch; the
    // pack is rebuilt (padding included) by the kernel's prologue, h/g are written before they are read
    a.barrier = reinterpret_cast<unsigned*>(p);
    a.dbg = reinterpret_cast<long long*>(p + 4096);
    a.block_part = reinterpret_cast<float*>(p + 8192);
    const size_t head = 8192 + sz.part_bytes;
    a.pack = reinterpret_cast<__nv_bfloat16*>(p + head);
    a.hbuf = reinterpret_cast<__nv_bfloat16*>(p + head + sz.pack_bytes);
    a.gbuf = reinterpret_cast<__nv_bfloat16*>(p + head + sz.pack_bytes + sz.h_bytes);
    cudaMemsetAsync(p, 0, head, s);
    // ensemble_loss entries of the member slice are written with plain stores (entries of other members
    // are left alone); loss_terms is accumulated with atomics
    cudaMemsetAsync(a.loss_terms, 0, sizeof(float) * (size_t)a.nsteps * 3, s);
    a.steps_per_epoch = (a.train_length + a.batch_size - 1) / a.batch_size;
    // dX reduction lengths of layers 1..L-1 pick the instantiation (13 k-steps cover widths up to 208)
    int ks = 0;
    for (int l = 1; l < a.L; ++l) ks = a.ly[l].ksteps_out > ks ? a.ly[l].ksteps_out : ks;
    const int inst = ks <= 13 ? 0 : 1;
    const void* kern = inst == 0 ? (const void*)k_train_fused<13> : (const void*)k_train_fused<16>;
    static int configured[2][64] = {{0}};
    if (dev < 64 && configured[inst][dev] < (int)sz.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sz.smem);
        if (e != cudaSuccess) {
            set_error("train_fused: %zu bytes of shared memory: %s", sz.smem, cudaGetErrorString(e));
            return MB_ECUDA;
        }
        configured[inst][dev] = (int)sz.smem;
    }
    void* args[] = {(void*)&a};
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)(a.ne * a.G)), dim3(kT), args, sz.smem, s);
    if (e != cudaSuccess) {
        set_error("k_train_fused launch (%d CTAs, %zu B smem): %s", a.ne * a.G, sz.smem, cudaGetErrorString(e));
        return MB_ECUDA;
    }
    return MB_OK;
}

}  // namespace mb
