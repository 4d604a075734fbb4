// Random-elite selection of PEModel.step on the device -- `np.random.choice(elite, B)`
// (mirl/agents/mbpo/pe_model.py:165-169), BIT-EXACT with NumPy's legacy global RandomState.
//
// NumPy draws `choice` as randint(0, n): one tempered MT19937 word per attempt, masked to the next
// power of two and rejected while > n-1 (random_bounded_uint64_fill, masked 32-bit branch).  At a
// million rows per rollout step that host loop (10-25 ms) is an order of magnitude slower than the
// rollout kernel it feeds, so the same stream is produced here in two stages:
//   1. k_mt_generate  (ONE CTA -- the recurrence is sequential block to block): regenerates the
//      624-word state block after block and streams the raw words to global memory.  It does
//      nothing else, because this CTA is issue/latency bound: tempering, rejection and compaction
//      in the same loop made it 4x slower (measured 1180 cycles per block against ~250).
//   2. k_draw_count / k_draw_scan / k_draw_scatter (whole GPU): temper + masked rejection of all
//      words, ordered compaction into the member index array, and the position of the B-th accepted
//      word, from which k_draw_finish leaves the generator state exactly where NumPy would be --
//      host code that keeps using np.random stays reproducible.
#include "common.cuh"
#include "pe_kernels.h"

namespace mb {

constexpr int kMtN = 624, kMtM = 397, kMtLag = kMtN - kMtM;      // lag 227
constexpr int kGenThreads = 256;            // thread t owns words t, t+227, t+454 (t < 227)
constexpr int kChunk = 2048;                // words per compaction CTA (256 threads x 8)

struct EliteTable { uint8_t v[64]; };

// workspace: [raw words nblocks*624 u32][chunk counts/offsets nchunks i64][jstar i64]
struct DrawPlan {
    int64_t nblocks, nwords, nchunks;
    uint32_t* raw;
    long long* offs;
    long long* jstar;
    size_t bytes;
};

static DrawPlan draw_plan(int64_t B, int n, void* ws) {
    DrawPlan p;
    uint32_t mask = (uint32_t)(n - 1);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    const double acc = (double)n / ((double)mask + 1.0);
    // expected attempts + 12 sigma of the negative-binomial count, + the (<= 624) words of the
    // input block that are already consumed, rounded up to whole blocks
    const double mean = (double)B / acc, sigma = sqrt((double)B * (1.0 - acc)) / acc;
    const int64_t need = (int64_t)(mean + 12.0 * sigma) + 2 * kMtN;
    p.nblocks = n == 1 ? 1 : (need + kMtN - 1) / kMtN + 1;
    p.nwords = p.nblocks * kMtN;
    p.nchunks = (p.nwords + kChunk - 1) / kChunk;
    size_t off = 0;
    p.raw = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ws) + off);
    off += (size_t)p.nwords * 4;
    off = (off + 15) & ~(size_t)15;
    p.offs = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(ws) + off);
    off += (size_t)(p.nchunks + 1) * 8;
    p.jstar = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(ws) + off);
    off += 16;
    p.bytes = off;
    return p;
}

size_t draw_workspace_bytes(int64_t B, int n) { return draw_plan(B, n, nullptr).bytes; }

__device__ __forceinline__ uint32_t mt_mix(uint32_t hi, uint32_t lo, uint32_t far) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// The recurrence new[k] = f(old[k], old[k+1], k < 227 ? old[k+397] : new[k-227]) chains words
// k, k+227, k+454 through ONE thread (each needs the previous one as its far operand), so a block
// regeneration needs no barrier between its three phases: every other operand is the old state,
// read from the shared-memory copy of the previous block.  Word 623 needs new[0]; its owner
// recomputes that single word instead of waiting for thread 0.  One barrier per 624 words.
// raw[0] = the input block (its first `pos` words are already consumed), raw[j] = j-th successor.
__global__ void __launch_bounds__(kGenThreads)
k_mt_generate(const uint32_t* __restrict__ state, uint32_t* __restrict__ raw, int64_t nblocks) {
    __shared__ uint32_t mt[2][kMtN + 1];
    const int t = threadIdx.x;
    for (int i = t; i < kMtN; i += kGenThreads) {
        const uint32_t s = state[i];
        mt[0][i] = s;
        raw[i] = s;
    }
    __syncthreads();
    const bool own = t < kMtLag;             // 227 owners
    const bool own2 = t < kMtN - 2 * kMtLag; // third slot: words 454..623 -> 170 owners
    const bool last = t == kMtN - 2 * kMtLag - 1;
    uint32_t x0 = 0, x1 = 0, x2 = 0;
    if (own) { x0 = mt[0][t]; x1 = mt[0][t + kMtLag]; }
    if (own2) x2 = mt[0][t + 2 * kMtLag];
    int cur = 0;
    for (int64_t j = 1; j < nblocks; ++j) {
        const uint32_t* o = mt[cur];
        uint32_t* nw = mt[cur ^ 1];
        uint32_t* g = raw + j * kMtN;
        if (own) {
            const uint32_t a = o[t + 1], b = o[t + kMtM], c = o[t + kMtLag + 1];
            uint32_t d = own2 ? o[last ? 0 : t + 2 * kMtLag + 1] : 0u;
            if (last) d = mt_mix(d, o[1], o[kMtM]);                      // new[0]
            x0 = mt_mix(x0, a, b);
            x1 = mt_mix(x1, c, x0);
            nw[t] = x0;
            nw[t + kMtLag] = x1;
            g[t] = x0;
            g[t + kMtLag] = x1;
            if (own2) {
                x2 = mt_mix(x2, d, x1);
                nw[t + 2 * kMtLag] = x2;
                g[t + 2 * kMtLag] = x2;
            }
        }
        cur ^= 1;
        __syncthreads();
    }
}

__device__ __forceinline__ bool draw_accept(const uint32_t* __restrict__ raw, int64_t g, int64_t nwords, int pos,
                                            uint32_t mask, uint32_t rng, uint32_t& v) {
    if (g >= nwords || g < pos) return false;
    v = mt_temper(raw[g]) & mask;
    return v <= rng;
}

// accepted words per chunk -> offs[chunk + 1]
__global__ void __launch_bounds__(256)
k_draw_count(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ state, int64_t nwords, uint32_t mask,
             uint32_t rng, long long* __restrict__ offs, long long* __restrict__ jstar) {
    const int pos = (int)state[kMtN];
    const int64_t g0 = (int64_t)blockIdx.x * kChunk + threadIdx.x;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kChunk / 256; ++k) {
        uint32_t v;
        c += draw_accept(raw, g0 + k * 256, nwords, pos, mask, rng, v) ? 1 : 0;
    }
    __shared__ int wsum[8];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < 8; ++i) s += wsum[i];
        offs[blockIdx.x + 1] = s;
        if (blockIdx.x == 0) { offs[0] = 0; *jstar = -1; }
    }
}

// in-place inclusive scan of offs[1..nchunks] (one CTA; nchunks is a few hundred per million rows)
__global__ void __launch_bounds__(1024)
k_draw_scan(long long* __restrict__ offs, int64_t nchunks) {
    __shared__ long long wtot[32];
    __shared__ long long carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nchunks; b0 += 1024) {
        const int64_t i = b0 + t;
        long long x = i < nchunks ? offs[i + 1] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += up;
        }
        if (lane == 31) wtot[w] = x;
        __syncthreads();
        if (w == 0) {
            long long y = wtot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long up = __shfl_up_sync(0xffffffffu, y, d);
                if (lane >= d) y += up;
            }
            wtot[lane] = y;
        }
        __syncthreads();
        const long long carry = carry_s;
        x += carry + (w > 0 ? wtot[w - 1] : 0);
        if (i < nchunks) offs[i + 1] = x;
        __syncthreads();
        if (t == 1023) carry_s = x;
        __syncthreads();
    }
}

// ordered compaction: the r-th accepted word (r < B) selects member_idx[r]; the word holding the
// B-th accept is recorded in jstar
__global__ void __launch_bounds__(256)
k_draw_scatter(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ state, int64_t nwords, uint32_t mask,
               uint32_t rng, EliteTable tab, const long long* __restrict__ offs, int64_t B,
               uint8_t* __restrict__ out, long long* __restrict__ jstar) {
    __shared__ uint8_t elite[64];
    __shared__ int wsum[kChunk / 32];        // accept count per 32-word group, in word order
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t < 64) elite[t] = tab.v[t];
    const int pos = (int)state[kMtN];
    const int64_t base = offs[blockIdx.x];
    if (base >= B) return;                   // uniform: everything here is beyond the draw
    const int64_t g0 = (int64_t)blockIdx.x * kChunk;
    uint32_t v[kChunk / 256];
    unsigned bal[kChunk / 256];
#pragma unroll
    for (int k = 0; k < kChunk / 256; ++k) {
        // group (k * 8 + w) covers words g0 + k*256 + w*32 + [0, 32)
        const bool a = draw_accept(raw, g0 + k * 256 + t, nwords, pos, mask, rng, v[k]);
        bal[k] = __ballot_sync(0xffffffffu, a);
        if (lane == 0) wsum[k * 8 + w] = __popc(bal[k]);
    }
    __syncthreads();
    // exclusive prefix over the 64 groups: two 32-wide warp scans per warp
    int lo = wsum[lane], hi = wsum[32 + lane];
    const int lo0 = lo, hi0 = hi;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int ul = __shfl_up_sync(0xffffffffu, lo, d), uh = __shfl_up_sync(0xffffffffu, hi, d);
        if (lane >= d) { lo += ul; hi += uh; }
    }
    const int lo_total = __shfl_sync(0xffffffffu, lo, 31);
    lo -= lo0;
    hi += lo_total - hi0;
#pragma unroll
    for (int k = 0; k < kChunk / 256; ++k) {
        const int grp = k * 8 + w;
        const int goff = __shfl_sync(0xffffffffu, grp < 32 ? lo : hi, grp & 31);
        const int64_t r = base + goff + __popc(bal[k] & ((1u << lane) - 1u));
        if (((bal[k] >> lane) & 1u) && r < B) {
            out[r] = elite[v[k]];
            if (r == B - 1) *jstar = g0 + k * 256 + t;
        }
    }
}

// state <- the block holding the B-th accepted word, pos just behind it; state[625..626] = number
// of members produced (== B unless the word budget ran out: then the state is the end of the budget)
__global__ void __launch_bounds__(256)
k_draw_finish(uint32_t* __restrict__ state, uint32_t* __restrict__ snap, const uint32_t* __restrict__ raw,
              int64_t nblocks, const long long* __restrict__ offs, int64_t nchunks,
              const long long* __restrict__ jstar, int64_t B) {
    const long long j = *jstar;
    const int64_t blk = j >= 0 ? j / kMtN : nblocks - 1;
    const uint32_t pos = j >= 0 ? (uint32_t)(j % kMtN) + 1u : (uint32_t)kMtN;
    const long long produced = j >= 0 ? (long long)B : offs[nchunks];
    __syncthreads();
    for (int i = threadIdx.x; i < kMtN; i += 256) {
        const uint32_t w = raw[blk * kMtN + i];
        state[i] = w;
        if (snap) snap[i] = w;
    }
    if (threadIdx.x == 0) {
        const uint32_t lo = (uint32_t)(produced & 0xffffffffll), hi = (uint32_t)(produced >> 32);
        state[kMtN] = pos; state[kMtN + 1] = lo; state[kMtN + 2] = hi;
        if (snap) { snap[kMtN] = pos; snap[kMtN + 1] = lo; snap[kMtN + 2] = hi; }
    }
}

__global__ void k_draw_single(uint8_t v, int64_t B, uint8_t* __restrict__ out, uint32_t* __restrict__ state,
                              uint32_t* __restrict__ snap) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            state[kMtN + 1] = (uint32_t)(B & 0xffffffffll);
            state[kMtN + 2] = (uint32_t)(B >> 32);
        }
        __syncthreads();
        if (snap)
            for (int i = threadIdx.x; i < kMtN + 3; i += blockDim.x) snap[i] = state[i];
    }
}

int launch_draw_members(uint32_t* state, uint32_t* snap, const uint8_t* elites, int n, int64_t B, uint8_t* out,
                        void* ws, cudaStream_t s) {
    if (n == 1) {                            // one elite: NumPy consumes no random words
        k_draw_single<<<148, 256, 0, s>>>(elites[0], B, out, state, snap);
        MB_CUDA_LAUNCH_CHECK("k_draw_single");
        return MB_OK;
    }
    EliteTable tab{};
    for (int i = 0; i < n; ++i) tab.v[i] = elites[i];
    const DrawPlan p = draw_plan(B, n, ws);
    const uint32_t rng = (uint32_t)(n - 1);
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    k_mt_generate<<<1, kGenThreads, 0, s>>>(state, p.raw, p.nblocks);
    MB_CUDA_LAUNCH_CHECK("k_mt_generate");
    k_draw_count<<<(unsigned)p.nchunks, 256, 0, s>>>(p.raw, state, p.nwords, mask, rng, p.offs, p.jstar);
    MB_CUDA_LAUNCH_CHECK("k_draw_count");
    k_draw_scan<<<1, 1024, 0, s>>>(p.offs, p.nchunks);
    MB_CUDA_LAUNCH_CHECK("k_draw_scan");
    k_draw_scatter<<<(unsigned)p.nchunks, 256, 0, s>>>(p.raw, state, p.nwords, mask, rng, tab, p.offs, B, out, p.jstar);
    MB_CUDA_LAUNCH_CHECK("k_draw_scatter");
    k_draw_finish<<<1, 256, 0, s>>>(state, snap, p.raw, p.nblocks, p.offs, p.nchunks, p.jstar, B);
    MB_CUDA_LAUNCH_CHECK("k_draw_finish");
    return MB_OK;
}

}  // namespace mb
