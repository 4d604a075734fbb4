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
// per-CTA wall-clock stamps (globaltimer is synchronised across SMs, clock64 is not): member 0, step 2
#define TL_ALL(base) do { if (threadIdx.x == 0 && step == 2 && blockIdx.x < a.G) { long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[(base) + blockIdx.x] = t_; } } while (0)
#else
#define TL_SET(i)
#define TL_ALL(base)
#endif

namespace mb {

namespace {

constexpr int kWarps = 8;        // MMA warps; lane 0 of the last one also issues the bulk copies of the ring
constexpr int kT = kWarps * 32;
constexpr int kThreads = kT;
constexpr int kIssuer = kWarps - 1;
constexpr int kRows = 16;        // batch rows per block (one MMA M tile)
constexpr int kChunk = 64;       // batch rows per streamed chunk of phase B (and the ring's stage size)
constexpr int kWChunk = 8 * kWarps;  // weight rows per streamed chunk of phase A: one dX output tile per warp (= 64)
constexpr int kNormMax = 512;    // cached normaliser columns: inputs (<= 256) + deltas (<= 256)
// Register budget: the kernel is one function, and whatever ptxas spills are kernel-lifetime values that get
// reloaded inside the hot loops of EVERY phase.  3 strips x 4 tiles of dW accumulators and 2 float4 per thread
// and sweep batch compile without a single spill (4 / 4: 88 bytes, and the whole step 4 % slower).
#ifndef MB_MAX_STRIPS
#define MB_MAX_STRIPS 3
#endif
#ifndef MB_SWEEP
#define MB_SWEEP 2
#endif
constexpr int kMaxStrips = MB_MAX_STRIPS;    // 16-row weight strips per phase-B job
constexpr int kSweep = MB_SWEEP;   // float4 per thread and batch of the Adam sweep

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
    }
    __syncthreads();
    // the bulk copies issued next read, through the async proxy, global data that other CTAs wrote with
    // ordinary stores
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
// activation and its derivative from one sigmoid: swish h = z s, h' = s (1 + z (1 - s)) = s + h (1 - s)
__device__ __forceinline__ float act_both(float z, int act, float& deriv) {
    if (act == MB_ACT_RELU) { deriv = z > 0.f ? 1.f : 0.f; return fmaxf(z, 0.f); }
    // sigmoid(z) = 1/2 + 1/2 tanh(z/2): one MUFU op instead of ex2 + rcp.  MUFU.TANH on sm_100a is far more
    // accurate than the 2^-11 the PTX manual guarantees (measured in the rollout kernel: same 4e-6 error)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    const float sg = fmaf(0.5f, t, 0.5f);
    const float h = z * sg;
    deriv = fmaf(h, 1.f - sg, sg);
    return h;
}
__device__ __forceinline__ float swish_both(float z, float& deriv) {     // act_both's swish branch, branch-free
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    const float sg = fmaf(0.5f, t, 0.5f);
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
// Layout: kWChunk-row chunks, each [hi rows x Ns][lo rows x Ns] contiguous -> a chunk is ONE bulk copy.
__device__ __forceinline__ int pack_row_off(int k, int half, int Kp, int Ns) {
    const int c = k / kWChunk;
    const int rows = min(kWChunk, Kp - c * kWChunk);
    return c * (kWChunk * Ns * 2) + half * rows * Ns + (k - c * kWChunk) * Ns;
}

struct StepScalars { float step_size, bc2_sqrt; };

}  // namespace

// ------------------------------------------------------------------------------------------------------
// Global layouts of the phase-A -> phase-B hand-over (bf16 elements):
//   g_l : [E][Bp/64][hi|lo][64][Ns]   a 64-row chunk of both halves is ONE bulk copy into the ring
//   h_l : [hi|lo][Kp/16 strips][E][Bp][16]   a job's strip x 64 rows is one 2 KB bulk copy; the two 16-byte
//         units of a row are swapped for rows with bit 2 set, which makes the transposed ldmatrix reads of
//         the 32-byte-stride tile bank-conflict free
__device__ __forceinline__ int64_t g_elem_off(const FusedLayer& ly, int e, int Bp, int b, int half) {
    return ly.g_off + ((((int64_t)e * (Bp >> 6) + (b >> 6)) * 2 + half) * kChunk + (b & 63)) * ly.Ns;
}
__device__ __forceinline__ int64_t h_elem_off(const FusedLayer& ly, int E, int e, int Bp, int b, int strip, int half) {
    return ly.h_off + ((((int64_t)half * (ly.Kp >> 4) + strip) * E + e) * Bp + b) * 16;
}


// ---- branch-free inner loops (the guarded generic loops below them keep every k-step behind a branch,
// which stops ptxas from overlapping a k-step's fragment loads with the previous k-step's MMAs) -----------
// forward: a chunk of NK k-steps (64 rows: 4), NTW n-tiles per warp (tiles warp, warp+8, ...)
template <int NK, int NTW>
__device__ __forceinline__ void fwd_chunk(float (&acc)[4][4], const __nv_bfloat16* Ah, const __nv_bfloat16* Al,
                                          const __nv_bfloat16* Wh, const __nv_bfloat16* Wl, int a_off, int b_off, int Ns,
                                          int warp) {
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
        uint32_t ah[4], al[4];
        ldsm_x4(ah, smem_u32(Ah + a_off + kk * 16));
        ldsm_x4(al, smem_u32(Al + a_off + kk * 16));
        uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int off = kk * 16 * Ns + b_off + (warp + kWarps * i) * 8;
            ldsm_x2_t(bh[i], smem_u32(Wh + off));
            ldsm_x2_t(bl[i], smem_u32(Wl + off));
        }
#pragma unroll
        for (int i = 0; i < NTW; ++i) mma3(acc[i], ah, al, bh[i], bl[i]);
    }
}
template <int NTW>
__device__ __forceinline__ void fwd_chunk_nk(int nk, float (&acc)[4][4], const __nv_bfloat16* Ah, const __nv_bfloat16* Al,
                                             const __nv_bfloat16* Wh, const __nv_bfloat16* Wl, int a_off, int b_off, int Ns,
                                             int warp) {
    switch (nk) {
    case 4: fwd_chunk<4, NTW>(acc, Ah, Al, Wh, Wl, a_off, b_off, Ns, warp); break;
    case 3: fwd_chunk<3, NTW>(acc, Ah, Al, Wh, Wl, a_off, b_off, Ns, warp); break;
    case 2: fwd_chunk<2, NTW>(acc, Ah, Al, Wh, Wl, a_off, b_off, Ns, warp); break;
    default: fwd_chunk<1, NTW>(acc, Ah, Al, Wh, Wl, a_off, b_off, Ns, warp); break;
    }
}
// phase B: a full 64-row chunk (4 groups of 16 batch rows), NS strips x NTW n-tiles per warp
template <int NS, int NTW>
__device__ __forceinline__ void dw_chunk4(float (&acc)[kMaxStrips][4][4], const __nv_bfloat16* Gh, const __nv_bfloat16* Gl,
                                          const __nv_bfloat16* Hh, const __nv_bfloat16* Hl, int g_off, int Ns, int warp,
                                          int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
        for (int i = 0; i < NTW; ++i) {
            const int off = kk * 16 * Ns + g_off + (warp + kWarps * i) * 8;
            ldsm_x2_t(bh[i], smem_u32(Gh + off));
            ldsm_x2_t(bl[i], smem_u32(Gl + off));
        }
        const int hb = kk * 16 + (lane & 7) + 8 * (lane >> 4);
        const int h_off = hb * 16 + ((((lane >> 3) & 1) ^ ((hb >> 2) & 1)) << 3);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            uint32_t ah[4], al[4];
            ldsm_x4_t(ah, smem_u32(Hh + s * (kChunk * 16) + h_off));
            ldsm_x4_t(al, smem_u32(Hl + s * (kChunk * 16) + h_off));
#pragma unroll
            for (int i = 0; i < NTW; ++i) mma3(acc[s][i], ah, al, bh[i], bl[i]);
        }
    }
}

// KS = compile-time bound of the dX reduction length in k-steps (its operand fragments live in registers)
// NT = n-tiles (8 columns) per consumer warp: widths up to 8 * kWarps * NT
template <int KS, int NT>
__global__ void __launch_bounds__(kThreads, 1)
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
    __shared__ int64_t s_rowidx[3][kRows];            // [step parity]: the step's first block (prefetched); [2]: further blocks
    __shared__ float s_norm[2 * kNormMax];            // [mean | std + eps] of [obs | act | delta] (constant over the launch)
    __shared__ float s_tgt[kRows][64 + 1];            // head targets of the block (normalised delta | reward) and the loss mask
    __shared__ float s_bound[2][64];                  // max / min log-std of this member
    __shared__ __align__(8) uint64_t s_bar[4];        // full[2] ("chunk landed"), empty[2] ("chunk consumed")
    const uint32_t full0 = smem_u32(&s_bar[0]), empty0 = smem_u32(&s_bar[2]);
    if (tid == 0) {
        mbar_init(full0, 1);
        mbar_init(full0 + 8, 1);
        mbar_init(empty0, kWarps);
        mbar_init(empty0 + 8, kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = tid; t < a.act_bytes / 4; t += kThreads) {      // padding columns must hold finite values
        reinterpret_cast<uint32_t*>(act0)[t] = 0u;
        reinterpret_cast<uint32_t*>(act1)[t] = 0u;
    }
    __syncthreads();
    // Ring protocol.  Use number u (0, 1, 2, ... over the whole kernel) occupies stage u & 1.  The issuer
    // (lane 0 of warp kIssuer, which has the fewest tiles) waits for the stage's previous use to be released
    // by all 8 warps and issues the bulk copy of use u + 1 before it consumes use u; a warp waits for the
    // copy to land, computes, and releases.  No CTA-wide barrier per chunk.
    uint32_t use = 0, use_p = 0;                      // uses consumed (every thread) / issued (issuer warp)
    auto producer_acquire = [&]() -> int {            // -> stage
        const int s = use_p & 1;
        if (use_p >= 2) mbar_wait(empty0 + 8 * s, ((use_p >> 1) & 1) ^ 1);
        ++use_p;
        return s;
    };
    auto consumer_wait = [&]() -> int {
        const int s = use & 1;
        mbar_wait(full0 + 8 * s, (use >> 1) & 1);
        ++use;
        return s;
    };
    auto consumer_release = [&](int s) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    };
    auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kT) : "memory"); };

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
        for (int t = tid; t < (k1 - k0) * npair; t += kThreads) {
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
    // Padding columns of the global copy of h_{l+1} (phase B's operand) beyond the tiles the forward epilogue
    // writes: column N is the bias column (exactly 1: it yields db), the rest 0 -- constants, written once for
    // the row blocks this CTA owns (phase B loads them into shared memory as part of its strips, so they must
    // be finite; rewriting them in every forward pass cost ~0.5 k cycles per pass)
    for (int blk = slot; blk < a.nblk_max; blk += G) {
        for (int l = 0; l + 1 < L; ++l) {
            const FusedLayer& nx = a.ly[l + 1];
            const int N = a.ly[l].N, c0 = (N + 7) & ~7, wpad = nx.Kp - c0;
            const int64_t half = (int64_t)(nx.Kp >> 4) * a.E * a.Bp * 16;
            for (int t = tid; t < kRows * wpad; t += kThreads) {
                const int i = t & (kRows - 1), c = c0 + (t >> 4), b = blk * kRows + i;
                const int unit = ((c >> 3) & 1) ^ ((b >> 2) & 1);
                const int64_t o = h_elem_off(nx, a.E, e, a.Bp, b, c >> 4, 0) + unit * 8 + (c & 7);
                hbuf[o] = __float2bfloat16(c == N ? 1.f : 0.f);
                hbuf[o + half] = __float2bfloat16(0.f);
            }
        }
    }
    bar_target += G;
    member_arrive(bar);
    member_wait(bar, bar_target);

    const int npass = 2 * L - 1;
    auto pass_layer = [&](int p) { return p < L ? p : 2 * L - 1 - p; };
    double b1t = a.beta1_pow0, b2t = a.beta2_pow0;    // beta^t as running products (thread 0)
    // batch of a step inside the epoch schedule (pe_model_trainer.py:240,268): first row and row count
    auto sched = [&](int step_, int64_t& j0_, int& B_) {
        const int64_t kb = (a.first_step + step_) % a.steps_per_epoch;
        j0_ = kb * a.batch_size;
        B_ = (int)min((int64_t)a.batch_size, a.train_length - j0_);
    };
    // data row of batch row b (-1 beyond the batch): the bootstrap table when there is one (pe_model_trainer.py:179)
    auto row_index = [&](int64_t j0_, int B_, int b) -> int64_t {
        if (b >= B_) return -1;
        return a.index ? a.index[(int64_t)e * a.index_stride + j0_ + b]
                       : (a.per_member_rows ? (int64_t)e * a.batch_size + b : (int64_t)b);
    };
    // the first block's row indices of step s live in s_rowidx[s & 1], fetched while step s - 1 is in phase B
    auto prefetch_rows = [&](int step_) {
        if (step_ >= a.nsteps || tid >= kRows) return;
        int64_t j0_;
        int B_;
        sched(step_, j0_, B_);
        s_rowidx[step_ & 1][tid] = row_index(j0_, B_, slot * kRows + tid);
    };
    prefetch_rows(0);
    if (a.norm != nullptr) {
        const int O = a.O, A = a.A, IN = O + A;
        for (int c = tid; c < IN + O; c += kThreads) {
            const float* mu = c < O ? a.norm + c : (c < IN ? a.norm + 2 * O + (c - O) : a.norm + 2 * O + 2 * A + (c - IN));
            const int w = c < O ? O : (c < IN ? A : O);
            s_norm[c] = mu[0];
            s_norm[kNormMax + c] = mu[w] + kNormEps;
        }
    }
    __syncthreads();

    for (int step = 0; step < a.nsteps; ++step) {
        int64_t j0;
        int B;
        sched(step, j0, B);
        const int nblk = (B + kRows - 1) / kRows;
        const float invB = 1.f / (float)B;
        float* const eloss = a.ensemble_loss + (int64_t)step * a.E;
        float* const lterms = a.loss_terms + (int64_t)step * 3;
        const int nchunksB = (nblk * kRows + kChunk - 1) / kChunk;

        // ================================ phase A (consumer warps) =================================
        TL_SET(0);
        TL_ALL(200);
        for (int blk = slot; blk < nblk; blk += G) {
            const int row0 = blk * kRows;
            // chunk stream over all layer passes of this block (forward 0..L-1, then dX L-1..1): ONE bulk copy
            // per chunk (the pack keeps a chunk's hi and lo halves adjacent), issued one chunk ahead
            int ip = 0, ic = 0;                           // next (pass, chunk) to issue
            auto issue_w = [&]() {
                if (warp != kIssuer || ip >= npass) return;
                const FusedLayer& ly = a.ly[pass_layer(ip)];
                const int st = producer_acquire();
                fence_proxy_async();
                if (lane == 0) {
                    const int rows = min(kWChunk, ly.Kp - ic * kWChunk);
                    const uint32_t bytes = (uint32_t)rows * ly.Ns * 4;
                    const uint32_t fb = full0 + 8 * st;
                    mbar_expect_tx(fb, bytes);
                    bulk_g2s(smem_u32(stage_ptr(st)),
                             pack + ly.pack_off + (int64_t)e * 2 * ly.Kp * ly.Ns + (int64_t)ic * (kWChunk * ly.Ns * 2), bytes, fb);
                }
                if (++ic * kWChunk >= ly.Kp) { ic = 0; ++ip; }
            };
            issue_w();
            // ---- row indices of the block (the step's first block: fetched a step ahead), then every global
            // operand of the staging in ONE round trip: the raw loads of the inputs and of the head operands are
            // issued before any of them is consumed (normaliser statistics sit in shared memory)
            int64_t* const ridx = s_rowidx[blk == slot ? (step & 1) : 2];
            if (blk != slot) {
                if (tid < kRows) ridx[tid] = row_index(j0, B, row0 + tid);
                consumer_sync();
            }
            {
                const FusedLayer& l0 = a.ly[0];
                const int As = l0.As, IN = l0.K, O = a.O, A = a.A, D = a.D;
                const bool nll = a.head_kind == FUSED_HEAD_NLL;
                __nv_bfloat16* ah = act0;
                __nv_bfloat16* al = ah + kRows * a.as_max;
                const int items1 = kRows * (As / 2), items2 = nll ? kRows * (D + 1) : 0;
                float bnd = 0.f;
                const bool has_bnd = nll && tid < 2 * 64 && (tid & 63) < D;
                if (has_bnd) bnd = __ldcg(a.params + ((tid >> 6) ? a.minls_off : a.maxls_off) + (int64_t)e * D + (tid & 63));
                for (int base = 0; base < max(items1, items2); base += 2 * kT) {
                    float xr[2][2], hr[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int t = base + tid + u * kT;
                        xr[u][0] = xr[u][1] = hr[u] = 0.f;
                        if (t < items1) {
                            const int i = t / (As / 2), j = (t % (As / 2)) * 2;
                            const int64_t r = ridx[i];
                            if (r >= 0) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const int c = j + q;
                                    if (c < O) xr[u][q] = __ldg(a.obs + r * O + c);
                                    else if (c < IN) xr[u][q] = __ldg(a.actn + r * A + (c - O));
                                }
                            }
                        }
                        if (t < items2) {
                            const int i = t / (D + 1), j = t - i * (D + 1);
                            const int64_t r = ridx[i];
                            if (r >= 0) {
                                if (j < O) hr[u] = __ldg(a.delta + r * O + j);
                                else if (j == O) hr[u] = __ldg(a.reward + r);
                                else hr[u] = a.cfg.ignore_terminal_state ? __ldg(a.done + r) : 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int t = base + tid + u * kT;
                        if (t < items1) {
                            const int i = t / (As / 2), j = (t % (As / 2)) * 2;
                            const bool live = ridx[i] >= 0;
                            float v[2] = {0.f, 0.f};
                            if (live) {
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const int c = j + q;
                                    // Normalizer.process (normalizer.py:17-18); s_norm = [mean | std + eps] per column
                                    if (c < IN) v[q] = a.norm ? (xr[u][q] - s_norm[c]) / s_norm[kNormMax + c] : xr[u][q];
                                    else if (c == IN) v[q] = 1.f;
                                }
                            }
                            uint32_t hi, lo;
                            split2(v[0], v[1], hi, lo);
                            *reinterpret_cast<uint32_t*>(ah + i * As + j) = hi;
                            *reinterpret_cast<uint32_t*>(al + i * As + j) = lo;
                            // h_0 for phase B (strip-major, swizzled 16-byte units): 4 bytes per thread
                            if (j < l0.Kp) {
                                const int b = row0 + i, unit = ((j >> 3) & 1) ^ ((b >> 2) & 1);
                                const int64_t o = h_elem_off(l0, a.E, e, a.Bp, b, j >> 4, 0) + unit * 8 + (j & 7);
                                *reinterpret_cast<uint32_t*>(hbuf + o) = hi;
                                *reinterpret_cast<uint32_t*>(hbuf + o + (int64_t)(l0.Kp >> 4) * a.E * a.Bp * 16) = lo;
                            }
                        }
                        // head operands of this block: fetched now, consumed five layers later
                        if (t < items2) {
                            const int i = t / (D + 1), j = t - i * (D + 1);
                            float v = 0.f;
                            if (ridx[i] >= 0) {
                                if (j < O) v = (hr[u] - s_norm[IN + j]) / s_norm[kNormMax + IN + j];
                                else if (j == O) v = hr[u];
                                else v = 1.f - hr[u];
                            }
                            s_tgt[i][j < D ? j : 64] = v;
                        }
                    }
                }
                if (has_bnd) s_bound[tid >> 6][tid & 63] = bnd;
            }
            consumer_sync();

            int cur = 0;                                  // act[cur] = input operand of the running pass
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
                const int nchunks = (ly.Kp + kWChunk - 1) / kWChunk;
                const int nt_total = fwd ? ly.nt_out : ly.nt_in;   // output n-tiles of this pass
                const int out_w = fwd ? ly.N : ly.K;               // real output columns
                // destination operand buffer of this pass
                __nv_bfloat16* Dh = act_ptr(cur ^ 1);
                __nv_bfloat16* Dl = Dh + kRows * a.as_max;
                const bool last_fwd = fwd && l == L - 1;
                // the produced operand: forward -> h_{l+1} (stride As of layer l+1); dX -> g_{l-1} (stride Ns of l-1)
                const FusedLayer& lo_ = a.ly[fwd ? (last_fwd ? l : l + 1) : l - 1];
                const int dstride = fwd ? lo_.As : lo_.Ns;
                float* const zrow = zbuf + (fwd ? l : l - 1) * kRows * a.z_stride;
                // global destination of this thread's two fragment rows (phase B's layouts), hoisted out of the tiles
                const int64_t h_strip_stride = (int64_t)a.E * a.Bp * 16, h_half = (int64_t)(lo_.Kp >> 4) * h_strip_stride;
                int64_t grow[2];
                int gsw[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int b = row0 + (lane >> 2) + 8 * h;
                    gsw[h] = (b >> 2) & 1;
                    grow[h] = fwd ? h_elem_off(lo_, a.E, e, a.Bp, b, 0, 0) + 2 * (lane & 3)
                                  : g_elem_off(lo_, e, a.Bp, b, 0) + 2 * (lane & 3);
                }

                // The epilogue of a tile writes the produced operand to shared memory (next pass) AND to
                // global memory in phase B's layout -- no separate copy pass.
                // Every decision that is uniform over the warp (last layer / full tile / activation kind) is taken once
                // per tile, and the bodies below are straight-line code: per-element branches on the activation kind
                // and on the column bounds kept the compiler from scheduling across them (9 % of the step's stall
                // samples were branch resolution) and made every epilogue a chain of tiny basic blocks.
                auto epilogue_tile = [&](int tile, const float (&c)[4]) {
                    // accumulator fragment: rows lane/4 and lane/4+8, columns n0 + 2*(lane%4), +1
                    const int n = tile * 8 + 2 * (lane & 3);
                    if (last_fwd) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = (lane >> 2) + 8 * h;
                            if (n < out_w) headbuf[i * a.head_stride + n] = c[2 * h];
                            if (n + 1 < out_w) headbuf[i * a.head_stride + n + 1] = c[2 * h + 1];
                        }
                        return;
                    }
                    const bool full = tile * 8 + 8 <= out_w;
                    float v[4];
                    if (fwd) {
                        // the backward pass needs act'(z), not z: keep that (one sigmoid serves both)
                        float d[4];
#ifdef MB_EXP_NOACT
#pragma unroll
                        for (int q = 0; q < 4; ++q) { v[q] = c[q]; d[q] = 1.f; }
#else
                        if (full && act_kind == MB_ACT_SWISH) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) v[q] = swish_both(c[q], d[q]);
                        } else if (full) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) { d[q] = c[q] > 0.f ? 1.f : 0.f; v[q] = fmaxf(c[q], 0.f); }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int col = n + (q & 1);
                                d[q] = 0.f;
                                if (col < out_w) v[q] = act_both(c[q], act_kind, d[q]);
                                else v[q] = col == out_w ? 1.f : 0.f;
                            }
                        }
#endif
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = (lane >> 2) + 8 * h;
                            uint32_t hi, lo;
                            *reinterpret_cast<float2*>(zrow + i * a.z_stride + n) = make_float2(d[2 * h], d[2 * h + 1]);
                            split2(v[2 * h], v[2 * h + 1], hi, lo);
#ifndef MB_EXP_NOGSTORE
                            const int64_t o = grow[h] + (tile >> 1) * h_strip_stride + (((tile & 1) ^ gsw[h]) << 3);
                            *reinterpret_cast<uint32_t*>(hbuf + o) = hi;
                            *reinterpret_cast<uint32_t*>(hbuf + o + h_half) = lo;
#endif
                            *reinterpret_cast<uint32_t*>(Dh + i * dstride + n) = hi;
                            *reinterpret_cast<uint32_t*>(Dl + i * dstride + n) = lo;
                        }
                    } else {
                        const float m0 = (full || n < out_w) ? 1.f : 0.f, m1 = (full || n + 1 < out_w) ? 1.f : 0.f;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = (lane >> 2) + 8 * h;
                            uint32_t hi, lo;
                            const float2 d = *reinterpret_cast<const float2*>(zrow + i * a.z_stride + n);
                            split2(c[2 * h] * d.x * m0, c[2 * h + 1] * d.y * m1, hi, lo);
                            const int64_t o = grow[h] + tile * 8;
                            *reinterpret_cast<uint32_t*>(gbuf + o) = hi;
                            *reinterpret_cast<uint32_t*>(gbuf + o + (int64_t)kChunk * lo_.Ns) = lo;
                            *reinterpret_cast<uint32_t*>(Dh + i * dstride + n) = hi;
                            *reinterpret_cast<uint32_t*>(Dl + i * dstride + n) = lo;
                        }
                    }
                };

                // Padding columns of the produced operand beyond the tiles the epilogue writes: the only one whose
                // VALUE matters is the bias column (column N of a forward operand must be exactly 1: it multiplies
                // the bias row of the next layer's pack and yields db in phase B); every other padding column only
                // ever meets zero weights, so any finite value will do -- the buffers are zeroed once at kernel start
                // and only ever hold finite bf16 values afterwards.  (Rewriting all of them every pass cost ~650
                // cycles per pass.)  When N is not a multiple of 8 the epilogue's last tile writes the 1 itself.
                // (the constant padding columns of the GLOBAL copy of h_{l+1} are written once, by the prologue)
                if (fwd && !last_fwd && (out_w & 7) == 0 && tid < kRows) {
                    Dh[tid * dstride + out_w] = __float2bfloat16(1.f);
                    Dl[tid * dstride + out_w] = __float2bfloat16(0.f);
                }
                TL_SET(60 + pass * 6 + 1);
                // per-lane ldmatrix offsets (elements) that do not depend on the chunk
                const int a_lane_off = ((lane & 7) + 8 * ((lane >> 3) & 1)) * astride + 8 * (lane >> 4);
                if (fwd) {
                    float acc[NT][4];
#pragma unroll
                    for (int i = 0; i < NT; ++i)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
                    const int b_lane_off = (lane & 15) * Ns;
                    for (int c = 0; c < nchunks; ++c) {
                        issue_w();
                        const int st = consumer_wait();
                        if (c == 0) TL_SET(60 + pass * 6 + 2);
                        const int rows = min(kWChunk, ly.Kp - c * kWChunk);
                        const __nv_bfloat16* Wh = stage_ptr(st);
                        const __nv_bfloat16* Wl = Wh + rows * Ns;
                        const int nk = rows / 16;
                        // reduction over the chunk's rows; warp owns n-tiles warp, warp+8, ...
                        const int ntw = nt_total > warp ? (nt_total - warp + kWarps - 1) / kWarps : 0;
                        static_assert(NT == 4, "the forward dispatch covers 0..4 n-tiles per warp");
                        const int ao = a_lane_off + c * kWChunk;
                        switch (ntw) {                               // warp-uniform; 0: nothing to do for this warp
                        case 4: fwd_chunk_nk<4>(nk, acc, Ah, Al, Wh, Wl, ao, b_lane_off, Ns, warp); break;
                        case 3: fwd_chunk_nk<3>(nk, acc, Ah, Al, Wh, Wl, ao, b_lane_off, Ns, warp); break;
                        case 2: fwd_chunk_nk<2>(nk, acc, Ah, Al, Wh, Wl, ao, b_lane_off, Ns, warp); break;
                        case 1: fwd_chunk_nk<1>(nk, acc, Ah, Al, Wh, Wl, ao, b_lane_off, Ns, warp); break;
                        default: break;
                        }
                        consumer_release(st);
                    }
                    TL_SET(60 + pass * 6 + 3);
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const int tile = warp + kWarps * i;
                        if (tile < nt_total) epilogue_tile(tile, acc[i]);
                    }
                } else {
                    // dX: every chunk needs the whole g_l operand (the reduction runs over its columns).  The hi
                    // fragments of all k-steps are loaded ONCE per pass and stay in registers (re-reading both
                    // halves per chunk made the pass shared-memory-bandwidth bound: 8 warps x 4 chunks x 13 KB);
                    // the lo halves are re-read -- keeping them too costs 52 more registers and spills
                    uint32_t gfh[KS][4];
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        if (kk < ly.ksteps_out) ldsm_x4(gfh[kk], smem_u32(Ah + a_lane_off + kk * 16));
                    }
                    for (int c = 0; c < nchunks; ++c) {
                        issue_w();
                        const int st = consumer_wait();
                        if (c == 0) TL_SET(60 + pass * 6 + 2);
                        const int rows = min(kWChunk, ly.Kp - c * kWChunk);
                        const __nv_bfloat16* Wh = stage_ptr(st);
                        const __nv_bfloat16* Wl = Wh + rows * Ns;
                        // chunk rows = output columns k; warp owns output tile 8c + warp; full reduction over n.
                        // six independent accumulator chains (3 terms x even/odd k-step): one tile per warp
                        // would otherwise be a 39-deep dependent MMA chain
                        const int tile = c * (kWChunk / 8) + warp;
                        if (tile < nt_total) {
                            float c6[6][4];
#pragma unroll
                            for (int i = 0; i < 6; ++i)
#pragma unroll
                                for (int q = 0; q < 4; ++q) c6[i][q] = 0.f;
                            const int b_lane_off = (warp * 8 + (lane & 7)) * Ns + 8 * ((lane >> 3) & 1);
                            if (ly.ksteps_out == KS) {
#pragma unroll
                                for (int kk = 0; kk < KS; ++kk) {
                                    uint32_t bh[2], bl[2], gl4[4];
                                    ldsm_x4(gl4, smem_u32(Al + a_lane_off + kk * 16));
                                    ldsm_x2(bh, smem_u32(Wh + b_lane_off + kk * 16));
                                    ldsm_x2(bl, smem_u32(Wl + b_lane_off + kk * 16));
                                    const int o = (kk & 1) * 3;
                                    mma16816(c6[o + 0], gl4, bh);
                                    mma16816(c6[o + 1], gfh[kk], bl);
                                    mma16816(c6[o + 2], gfh[kk], bh);
                                }
                            } else {
#pragma unroll
                                for (int kk = 0; kk < KS; ++kk) {
                                    if (kk < ly.ksteps_out) {
                                        uint32_t bh[2], bl[2], gl4[4];
                                        ldsm_x4(gl4, smem_u32(Al + a_lane_off + kk * 16));
                                        ldsm_x2(bh, smem_u32(Wh + b_lane_off + kk * 16));
                                        ldsm_x2(bl, smem_u32(Wl + b_lane_off + kk * 16));
                                        const int o = (kk & 1) * 3;
                                        mma16816(c6[o + 0], gl4, bh);
                                        mma16816(c6[o + 1], gfh[kk], bl);
                                        mma16816(c6[o + 2], gfh[kk], bh);
                                    }
                                }
                            }
                            consumer_release(st);              // the stage is free before the epilogue maths
                            float cacc[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                cacc[q] = ((c6[0][q] + c6[3][q]) + (c6[1][q] + c6[4][q])) + (c6[2][q] + c6[5][q]);
                            epilogue_tile(tile, cacc);
                        } else {
                            consumer_release(st);
                        }
                    }
                    TL_SET(60 + pass * 6 + 3);
                }
                TL_SET(60 + pass * 6 + 4);
                consumer_sync();
                TL_SET(60 + pass * 6 + 5);

                if (last_fwd) {
                    // ---- loss head: writes g_{L-1} (bf16 hi/lo, stride Ns) into the destination buffer and
                    // to global memory
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
                    consumer_sync();
                    if (a.head_kind == FUSED_HEAD_NLL) {
                        for (int t = tid; t < kRows * D; t += kT) {
                            const int i = t / D, j = t - i * D;
                            if (ridx[i] < 0) continue;
                            const float mu = headbuf[i * a.head_stride + j], raw = headbuf[i * a.head_stride + D + j];
                            const float mx = s_bound[0][j], mn = s_bound[1][j];
                            const float aa = mx - raw;
                            const float ls1 = mx - softplusf_(aa);
                            const float cc = ls1 - mn;
                            const float ls = mn + softplusf_(cc);
                            const float tgt = s_tgt[i][j];
                            const float coef = j < O ? 1.f : a.cfg.reward_coef;
                            const float mask = s_tgt[i][64];
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
                            float diff = q - a.target[(int64_t)b * N + j];
                            if (a.clip_error > 0.f) diff = fminf(fmaxf(diff, -a.clip_error), a.clip_error);
                            part += diff * diff * wgt;
                            const float g = 2.f * diff * wgt;
                            const __nv_bfloat16 gh = __float2bfloat16(g);
                            Dh[i * gs + j] = gh;
                            Dl[i * gs + j] = __float2bfloat16(g - __bfloat162float(gh));
                        }
                    }
                    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if (lane == 0) s_part[warp] = part;
                    consumer_sync();
                    // g_{L-1} to global (phase B), 16-byte pieces
                    {
                        const int ppr = gs / 8;
                        for (int p = tid; p < kRows * ppr; p += kT) {
                            const int i = p / ppr, q = p - i * ppr;
                            const int64_t o = g_elem_off(ly, e, a.Bp, row0 + i, 0) + q * 8;
                            *reinterpret_cast<uint4*>(gbuf + o) = *reinterpret_cast<const uint4*>(Dh + i * gs + q * 8);
                            *reinterpret_cast<uint4*>(gbuf + o + (int64_t)kChunk * gs) = *reinterpret_cast<const uint4*>(Dl + i * gs + q * 8);
                        }
                    }
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
                    consumer_sync();
                }
                cur ^= 1;
            }
        }
        prefetch_rows(step + 1);
        TL_SET(1);
        TL_ALL(230);
        bar_target += G;
        member_arrive(bar);
        member_wait(bar, bar_target);
        TL_SET(2);
        TL_ALL(290);

        // ================================ phase B (consumer warps) =================================
        if (tid == 0) {
            b1t *= (double)a.cfg.beta1;
            b2t *= (double)a.cfg.beta2;
            s_sc.step_size = (float)((double)a.cfg.lr / (1.0 - b1t));
            s_sc.bc2_sqrt = (float)sqrt(1.0 - b2t);
        }
        consumer_sync();
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
            const int Ns = ly.Ns, ns = jb.nstrips;
            float acc[kMaxStrips][NT][4];
#pragma unroll
            for (int s = 0; s < kMaxStrips; ++s)
#pragma unroll
                for (int i = 0; i < NT; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[s][i][q] = 0.f;
            const int nt_total = ly.nt_out;
            // g chunk: one bulk copy (both halves); h tiles [half][strip][64][16]: one 2 KB copy per strip and half
            auto issue_b = [&](int c) {
                if (warp != kIssuer) return;
                const int st = producer_acquire();
                fence_proxy_async();        // the ring doubles as epilogue scratch (generic-proxy writes)
                const uint32_t fb = full0 + 8 * st;
                const uint32_t gbytes = (uint32_t)kChunk * Ns * 4;
                if (lane == 0) {
                    mbar_expect_tx(fb, gbytes + (uint32_t)ns * 2 * (kChunk * 16 * 2));
                    bulk_g2s(smem_u32(stage_ptr(st)), gbuf + g_elem_off(ly, e, a.Bp, c * kChunk, 0), gbytes, fb);
                }
                __syncwarp();
                if (lane < 2 * ns) {
                    const int half = lane / ns, sl = lane - half * ns;
                    bulk_g2s(smem_u32(act_ptr(st) + (half * ns + sl) * (kChunk * 16)),
                             hbuf + h_elem_off(ly, a.E, e, a.Bp, c * kChunk, jb.strip0 + sl, half), kChunk * 16 * 2, fb);
                }
            };
            const int g_lane_off = (lane & 15) * Ns;
            issue_b(0);
            for (int c = 0; c < nchunksB; ++c) {
                if (c + 1 < nchunksB) issue_b(c + 1);
                const int st = consumer_wait();
                const __nv_bfloat16* Gh = stage_ptr(st);
                const __nv_bfloat16* Gl = Gh + kChunk * Ns;
                const __nv_bfloat16* Hh = act_ptr(st);        // [hi|lo][strip][64 rows][16], swizzled units
                const __nv_bfloat16* Hl = Hh + ns * (kChunk * 16);
                const int groups = min(kChunk / 16, nblk - c * (kChunk / 16));
                const int ntw = nt_total > warp ? (nt_total - warp + kWarps - 1) / kWarps : 0;
                bool fast = false;
                if constexpr (NT == 4) {
                    if (groups == 4 && (ntw >= 3 || ntw <= 1)) {
                        fast = true;
                        if (ntw == 4) {
#if MB_MAX_STRIPS >= 4
                            if (ns == 4) dw_chunk4<4, 4>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else
#endif
                            if (ns == 3) dw_chunk4<3, 4>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else if (ns == 2) dw_chunk4<2, 4>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else dw_chunk4<1, 4>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                        } else if (ntw == 3) {
#if MB_MAX_STRIPS >= 4
                            if (ns == 4) dw_chunk4<4, 3>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else
#endif
                            if (ns == 3) dw_chunk4<3, 3>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else if (ns == 2) dw_chunk4<2, 3>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else dw_chunk4<1, 3>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                        } else if (ntw == 1) {                  // narrow layers (the 36-wide head): one tile per warp
#if MB_MAX_STRIPS >= 4
                            if (ns == 4) dw_chunk4<4, 1>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else
#endif
                            if (ns == 3) dw_chunk4<3, 1>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else if (ns == 2) dw_chunk4<2, 1>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                            else dw_chunk4<1, 1>(acc, Gh, Gl, Hh, Hl, g_lane_off, Ns, warp, lane);
                        }                                       // ntw == 0: nothing to do for this warp
                    }
                }
                if (!fast)
#pragma unroll
                for (int kk = 0; kk < kChunk / 16; ++kk) {
                    if (kk >= groups) break;
                    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const int tile = warp + kWarps * i;
                        if (tile < nt_total) {
                            const int off = kk * 16 * Ns + g_lane_off + tile * 8;
                            ldsm_x2_t(bh[i], smem_u32(Gh + off));
                            ldsm_x2_t(bl[i], smem_u32(Gl + off));
                        }
                    }
                    // A fragment (transposed load): matrix mi -> k half (mi & 1), row block (mi >> 1)
                    const int hb = kk * 16 + (lane & 7) + 8 * (lane >> 4);
                    const int h_off = hb * 16 + ((((lane >> 3) & 1) ^ ((hb >> 2) & 1)) << 3);
#pragma unroll
                    for (int s = 0; s < kMaxStrips; ++s) {
                        if (s < ns) {
                            uint32_t ah[4], al[4];
                            ldsm_x4_t(ah, smem_u32(Hh + s * (kChunk * 16) + h_off));
                            ldsm_x4_t(al, smem_u32(Hl + s * (kChunk * 16) + h_off));
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const int tile = warp + kWarps * i;
                                if (tile < nt_total) mma3(acc[s][i], ah, al, bh[i], bl[i]);
                            }
                        }
                    }
                }
                consumer_release(st);
            }
            TL_SET(40 + j);
            consumer_sync();                                  // every warp is done reading the ring
            // ---- epilogue: the accumulators go through shared memory (the chunk ring is idle now) so that
            // weight decay + Adam + the pack refresh run as a row-contiguous, 16-byte-vectorised sweep over
            // parameters, moments and pack (fragment-order accesses were 32-byte scattered: the sweep was
            // bound by memory transactions, not bytes)
            float* dW = reinterpret_cast<float*>(stage0);          // [16 * ns][Ns] fp32
#pragma unroll
            for (int s = 0; s < kMaxStrips; ++s) {
                if (s >= ns) continue;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
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
            consumer_sync();
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
                if (vec4) {
                    // Flat sweep, kSweep float4 per thread and batch: all loads of a batch are issued before
                    // the first store (a row-by-row loop is one memory round trip per iteration -- the compiler
                    // may not move loads above stores that could alias -- and the sweep was latency bound).
                    // Parameters and both moments share one blob layout, so one offset addresses all three.
                    const int n4 = N >> 2, total = krows * n4;
                    const int64_t wbase = ly.w_off + (int64_t)e * ly.K * N, bbase = ly.b_off + (int64_t)e * N;
                    for (int f0 = tid; f0 < total; f0 += kSweep * kT) {
                        float4 w[kSweep], m1[kSweep], m2[kSweep];
                        int64_t off[kSweep];
                        int rr[kSweep], nn[kSweep];
#pragma unroll
                        for (int u = 0; u < kSweep; ++u) {
                            const int f = f0 + u * kT;
                            rr[u] = -1;
                            if (f < total) {
                                const int r = f / n4, n = (f - r * n4) * 4;
                                const int k = jb.strip0 * 16 + r;
                                rr[u] = r;
                                nn[u] = n;
                                off[u] = k < ly.K ? wbase + (int64_t)k * N + n : bbase + n;
                                w[u] = *reinterpret_cast<const float4*>(a.params + off[u]);
                                m1[u] = *reinterpret_cast<const float4*>(a.exp_avg + off[u]);
                                m2[u] = *reinterpret_cast<const float4*>(a.exp_avg_sq + off[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kSweep; ++u) {
                            if (rr[u] < 0) continue;
                            const int k = jb.strip0 * 16 + rr[u], n = nn[u];
                            const float wd = k < ly.K ? ly.wd : 0.f;
                            const float4 g = *reinterpret_cast<const float4*>(dW + rr[u] * Ns + n);
                            const float4 wv = w[u];
                            l2_part += 0.5f * wd * ((wv.x * wv.x + wv.y * wv.y) + (wv.z * wv.z + wv.w * wv.w));
                            float4 nw;
                            nw.x = adam(wv.x, m1[u].x, m2[u].x, fmaf(wd, wv.x, g.x));
                            nw.y = adam(wv.y, m1[u].y, m2[u].y, fmaf(wd, wv.y, g.y));
                            nw.z = adam(wv.z, m1[u].z, m2[u].z, fmaf(wd, wv.z, g.z));
                            nw.w = adam(wv.w, m1[u].w, m2[u].w, fmaf(wd, wv.w, g.w));
                            *reinterpret_cast<float4*>(a.params + off[u]) = nw;
                            *reinterpret_cast<float4*>(a.exp_avg + off[u]) = m1[u];
                            *reinterpret_cast<float4*>(a.exp_avg_sq + off[u]) = m2[u];
                            uint2 hi, lo;
                            split2(nw.x, nw.y, hi.x, lo.x);
                            split2(nw.z, nw.w, hi.y, lo.y);
                            *reinterpret_cast<uint2*>(pm + pack_row_off(k, 0, ly.Kp, Ns) + n) = hi;
                            *reinterpret_cast<uint2*>(pm + pack_row_off(k, 1, ly.Kp, Ns) + n) = lo;
                        }
                    }
                } else {
                    for (int r = warp; r < krows; r += kWarps) {                // one weight row per warp
                        const int k = jb.strip0 * 16 + r;
                        const bool isw = k < ly.K;
                        float* rp = isw ? Wp + (int64_t)k * N : Bq;
                        float* r1 = isw ? W1 + (int64_t)k * N : B1;
                        float* r2 = isw ? W2 + (int64_t)k * N : B2;
                        const float wd = isw ? ly.wd : 0.f;
                        __nv_bfloat16* ph = pm + pack_row_off(k, 0, ly.Kp, Ns);
                        __nv_bfloat16* pl = pm + pack_row_off(k, 1, ly.Kp, Ns);
                        const float* dr = dW + r * Ns;
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
            consumer_sync();                                  // scratch no longer needed: the next job may refill the ring
        }
        // L2 term of the loss (mlp.py:170-178), from the pre-update weights
        for (int o = 16; o; o >>= 1) l2_part += __shfl_xor_sync(0xffffffffu, l2_part, o);
        if (lane == 0 && l2_part != 0.f) atomicAdd(lterms + 1, l2_part);
        // per-member loss and log-std bounds (pe_model.py:62-69,105-108): the row blocks' partials are
        // summed in a fixed order; Adam on 2 x D scalars per member + the bound term of the loss
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
            consumer_sync();
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
        TL_ALL(260);
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
        if (d.sizes[l] > 8 * kWarps * 4) return false;               // widths up to 256
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
        h_elems += (int64_t)2 * E * a.Bp * ly.Kp;            // [hi|lo][Kp/16][E][Bp][16]
        ly.g_off = g_elems;
        g_elems += (int64_t)2 * E * a.Bp * ly.Ns;            // [E][Bp/64][hi|lo][64][Ns]
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
    const int htile_b = 2 * kMaxStrips * kChunk * 16 * 2;        // [hi|lo][strip][64][16]
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
    const void* kern = inst == 0 ? (const void*)k_train_fused<13, 4> : (const void*)k_train_fused<16, 4>;
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
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)(a.ne * a.G)), dim3(kThreads), args, sz.smem, s);
    if (e != cudaSuccess) {
        set_error("k_train_fused launch (%d CTAs, %zu B smem): %s", a.ne * a.G, sz.smem, cudaGetErrorString(e));
        return MB_ECUDA;
    }
    return MB_OK;
}

}  // namespace mb
