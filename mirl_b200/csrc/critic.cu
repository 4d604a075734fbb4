// Critic-ensemble kernels (SIMT fp32): plain ensemble forward, REDQ random-subset target and
// TQC truncated-quantile target.
//   EnsembleQValue.value      mirl/values/ensemble_value.py:65-107  (+ sample_from_ensemble :10-34)
//   SACAgent.compute_q_target mirl/agents/sac_agent.py:73-86
//   TQCQValue.value           mirl/values/distributional_value.py:32-93
//   TQCAgent.compute_q_target mirl/agents/tqc_agent.py:41-57
// Unlike the reference (which evaluates all E critics and indexes the result), the batch-wise
// REDQ target evaluates only the drawn members.
#include "common.cuh"
#include "simt_mlp.cuh"
#include "layer_gemm.cuh"
#include "tc_gemm.h"
#include <stdlib.h>
#include "pe_kernels.h"

namespace mb {

constexpr int kMaxSub = 32;
struct MemberList { int n; int id[kMaxSub]; };

template <class Tile>
__device__ __forceinline__ void stage_features(const float* __restrict__ obs,
                                               const float* __restrict__ act, int O, int A,
                                               int64_t row0, int cnt, float* x0) {
    const int IN = O + A, IN4 = (IN + 3) & ~3;
    for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads) {
        const int i = t / IN4, j = t - i * IN4;
        float v = 0.f;
        if (i < cnt && j < IN) v = j < O ? obs[(row0 + i) * O + j] : act[(row0 + i) * A + (j - O)];
        x0[i * Tile::AS + j] = v;
    }
}

// out[E][B][N]; grid = (row tiles, E)
template <class Tile>
__global__ void __launch_bounds__(kSimtThreads, 1)
k_critic_forward(mb_mlp_desc d, const float* __restrict__ params, const float* __restrict__ obs,
                 const float* __restrict__ act, int O, int A, int64_t B, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    const int e = blockIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * Tile::TM;
    const int cnt = (int)min((int64_t)Tile::TM, B - row0);
    const int N = d.sizes[d.num_layers];
    stage_features<Tile>(obs, act, O, A, row0, cnt, buf0);
    __syncthreads();
    const float* res = Tile::forward(d, params, e, buf0, buf1, wbuf, nullptr, cnt);
    for (int t = threadIdx.x; t < cnt * N; t += kSimtThreads) {
        const int i = t / N, j = t - i * N;
        out[((int64_t)e * B + row0 + i) * N + j] = res[i * Tile::AS + j];
    }
}

// REDQ target.  ml = members to evaluate (the batch-wise draw, or all members when the subset
// is per row / absent).  row_members [n_sub][B] selects, per row, which evaluated members count.
template <class Tile>
__global__ void __launch_bounds__(kSimtThreads, 1)
k_critic_redq(mb_mlp_desc d, const float* __restrict__ params, const float* __restrict__ obs,
              const float* __restrict__ act, int O, int A, int64_t B, MemberList ml,
              const uint8_t* __restrict__ row_members, int n_sub, int mode, int td,
              const float* __restrict__ reward, const float* __restrict__ terminal,
              const float* __restrict__ log_prob, float alpha, float discount,
              float* __restrict__ value) {
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    const int IN4 = (O + A + 3) & ~3;
    float* x0 = wbuf + 2 * Tile::kWFloats;     // [TM][IN4]
    float* qacc = x0 + Tile::TM * IN4;         // [TM]
    const int64_t row0 = (int64_t)blockIdx.x * Tile::TM;
    const int cnt = (int)min((int64_t)Tile::TM, B - row0);
    stage_features<Tile>(obs, act, O, A, row0, cnt, buf0);
    __syncthreads();
    for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads)
        x0[t] = buf0[(t / IN4) * Tile::AS + (t % IN4)];
    if (threadIdx.x < Tile::TM)
        qacc[threadIdx.x] = mode == MB_REDUCE_MIN ? INFINITY : mode == MB_REDUCE_MAX ? -INFINITY : 0.f;
    __syncthreads();
    for (int k = 0; k < ml.n; ++k) {
        const int e = ml.id[k];
        if (k > 0) {
            for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads)
                buf0[(t / IN4) * Tile::AS + (t % IN4)] = x0[t];
            __syncthreads();
        }
        const float* res = Tile::forward(d, params, e, buf0, buf1, wbuf, nullptr, cnt);
        if (threadIdx.x < cnt) {
            const int i = threadIdx.x;
            bool use = true;
            if (row_members) {
                use = false;
                for (int s = 0; s < n_sub; ++s) use = use || (row_members[(int64_t)s * B + row0 + i] == e);
            }
            if (use) {
                const float q = res[i * Tile::AS];
                qacc[i] = mode == MB_REDUCE_MIN ? fminf(qacc[i], q)
                        : mode == MB_REDUCE_MAX ? fmaxf(qacc[i], q) : qacc[i] + q;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < cnt) {
        const int64_t r = row0 + threadIdx.x;
        float v = qacc[threadIdx.x];
        if (mode == MB_REDUCE_MEAN) v /= (float)n_sub;
        if (td) v = reward[r] + (1.f - terminal[r]) * discount * (v - alpha * log_prob[r]);
        value[r] = v;
    }
}

// TQC target: all members, pool E*Q atoms per row, ascending sort, drop the n_drop largest.
template <class Tile>
__global__ void __launch_bounds__(kSimtThreads, 1)
k_critic_tqc(mb_mlp_desc d, const float* __restrict__ params, const float* __restrict__ obs,
             const float* __restrict__ act, int O, int A, int64_t B, int n_drop, int npad, int td,
             const float* __restrict__ reward, const float* __restrict__ terminal,
             const float* __restrict__ log_prob, float alpha, float discount,
             float* __restrict__ value, float* __restrict__ atoms) {
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    const int IN4 = (O + A + 3) & ~3;
    float* x0 = wbuf + 2 * Tile::kWFloats;     // [TM][IN4]
    float* pool = x0 + Tile::TM * IN4;         // [TM][npad]
    const int E = d.ensemble_size, Q = d.sizes[d.num_layers], NA = E * Q, keep = NA - n_drop;
    const int64_t row0 = (int64_t)blockIdx.x * Tile::TM;
    const int cnt = (int)min((int64_t)Tile::TM, B - row0);
    stage_features<Tile>(obs, act, O, A, row0, cnt, buf0);
    for (int t = threadIdx.x; t < Tile::TM * npad; t += kSimtThreads) pool[t] = INFINITY;
    __syncthreads();
    for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads)
        x0[t] = buf0[(t / IN4) * Tile::AS + (t % IN4)];
    __syncthreads();
    for (int e = 0; e < E; ++e) {
        if (e > 0) {
            for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads)
                buf0[(t / IN4) * Tile::AS + (t % IN4)] = x0[t];
            __syncthreads();
        }
        const float* res = Tile::forward(d, params, e, buf0, buf1, wbuf, nullptr, cnt);
        // member-major within a row: ensemble_atoms.transpose(-3,-2).reshape(B, E*Q)
        for (int t = threadIdx.x; t < cnt * Q; t += kSimtThreads) {
            const int i = t / Q, q = t - i * Q;
            pool[i * npad + e * Q + q] = res[i * Tile::AS + q];
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < cnt; i += kSimtThreads / 32) {
        float* row = pool + i * npad;
        if (n_drop > 0) {
            // bitonic sort of npad (power of two) values, +inf padded, one warp per row
            for (int k = 2; k <= npad; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < npad / 2; t += 32) {
                        const int lo = ((t / j) * 2 * j) + (t % j);
                        const int hi = lo + j;
                        const bool up = ((lo & k) == 0);
                        const float a = row[lo], b = row[hi];
                        if ((a > b) == up) { row[lo] = b; row[hi] = a; }
                    }
                    __syncwarp();
                }
            }
        }
        const int64_t r = row0 + i;
        float sum = 0.f;
        for (int t = lane; t < keep; t += 32) sum += row[t];
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0 && value) value[r] = sum / (float)keep;
        if (atoms) {
            float rw = 0.f, sc = 1.f, sh = 0.f;
            if (td) { rw = reward[r]; sc = (1.f - terminal[r]) * discount; sh = alpha * log_prob[r]; }
            for (int t = lane; t < keep; t += 32)
                atoms[r * keep + t] = td ? rw + sc * (row[t] - sh) : row[t];
        }
    }
}


// ------------------------------------------------------------------------------------------
// Small-batch path: one launch per layer (layer_gemm.cuh) + a reduce / sort kernel.
// ------------------------------------------------------------------------------------------
constexpr int64_t kLayeredMaxRows = 1 << 20;      // rows x members up to which the layered path is used

size_t critic_workspace_bytes(const mb_mlp_desc& d, int64_t B) {
    int w = 0;
    for (int l = 0; l <= d.num_layers; ++l) w = d.sizes[l] > w ? d.sizes[l] : w;
    return (size_t)(2 * (int64_t)d.ensemble_size * B * w + (int64_t)d.ensemble_size * B * d.sizes[d.num_layers]) * 4 + 256;
}

// MB_CRITIC_GEMM=simt keeps every layer on the fp32 SIMT GEMM (A/B timing)
static bool critic_gemm_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MB_CRITIC_GEMM");
        v = (e != nullptr && e[0] == 's') ? 0 : 1;
    }
    return v == 1;
}

// Runs the stack for `nmem` members; returns the final layer's output [nmem][B][N_out] (in ws
// unless `final_out` is given).
// `layers` < num_layers stops after that many layers and returns the (activated) hidden rows
// [nmem][B][sizes[layers]] -- the caller then folds the remaining narrow output layer into its
// own reduction kernel instead of paying a GEMM launch for it.
static int layered_forward(const mb_mlp_desc& d, const float* params, const float* obs, const float* act,
                           int O, int A, int64_t B, const int* members, int nmem, float* ws, float* final_out,
                           float** result, cudaStream_t s, int layers = -1) {
    if (layers < 0) layers = d.num_layers;
    int w = 0;
    for (int l = 0; l <= d.num_layers; ++l) w = d.sizes[l] > w ? d.sizes[l] : w;
    float* buf[2] = {ws, ws + (int64_t)d.ensemble_size * B * w};
    float* last = final_out ? final_out : ws + 2 * (int64_t)d.ensemble_size * B * w;
    ConcatIn ci{obs, act, nullptr, O, A, 0};
    for (int l = 0; l < layers; ++l) {
        LayerArgs a;
        a.B = (int)B; a.K = d.sizes[l]; a.N = d.sizes[l + 1];
        a.A = l == 0 ? nullptr : buf[(l - 1) & 1];
        a.a_slot_stride = (int64_t)B * a.K;
        a.W = params + w_offset(d, l);
        a.bias = params + b_offset(d, l);
        a.out = l + 1 == d.num_layers ? last : buf[l & 1];
        a.a_act = -1;
        a.out_act = l + 1 < d.num_layers ? d.activation : -1;
        a.slot0 = 0;
        a.nmem = nmem;
        for (int i = 0; i < nmem; ++i) a.member[i] = members[i];
        if (l > 0 && critic_gemm_tc()) {
            // hidden / output layers on the tensor cores (bf16x3, tc_gemm.cu); layer 0 keeps the
            // SIMT kernel: its 23-wide input is the fused [obs | act] concatenation
            TgBatch tb{};
            tb.nops = 1; tb.nmem = nmem; tb.slot0 = 0;
            for (int i = 0; i < nmem; ++i) tb.member[i] = members[i];
            TgOp& o = tb.op[0];
            o.mode = TG_FWD;
            o.M = (int)B; o.N = a.N; o.R = a.K;
            o.a = a.A; o.lda = a.K; o.a_ss = a.a_slot_stride; o.a_act = -1; o.a_ones_row = -1;
            o.b = a.W; o.ldb = a.N; o.b_ms = (int64_t)a.K * a.N;
            o.bias = a.bias; o.bias_ms = a.N;
            o.out_act = a.out_act;
            o.out = a.out; o.ldo = a.N; o.out_ss = (int64_t)B * a.N;
            int rc = launch_tc_gemm(tb, s);
            if (rc) return rc;
            continue;
        }
        int rc = launch_layer(a, l == 0 ? &ci : nullptr, s);
        if (rc) return rc;
    }
    *result = layers == d.num_layers ? last : buf[(layers - 1) & 1];
    return MB_OK;
}

// The same reduction with the critics' single-output last layer folded in: one warp per row,
// q_k = h[k][r][:] . W_last[e_k] + b_last[e_k] (ensemble_value.py:65-107 after the hidden stack).
__global__ void __launch_bounds__(256)
k_redq_reduce_dot(const float* __restrict__ h, int K, const float* __restrict__ Wl, const float* __restrict__ bl,
                  int nmem, MemberList ml, int64_t B, const uint8_t* __restrict__ row_members, int n_sub, int mode,
                  int td, const float* __restrict__ reward, const float* __restrict__ terminal,
                  const float* __restrict__ log_prob, float alpha, float discount, float* __restrict__ value) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= B) return;
    float acc = mode == MB_REDUCE_MIN ? INFINITY : mode == MB_REDUCE_MAX ? -INFINITY : 0.f;
    for (int k = 0; k < nmem; ++k) {
        bool use = true;
        if (row_members) {
            use = false;
            for (int sidx = 0; sidx < n_sub; ++sidx) use = use || (row_members[(int64_t)sidx * B + r] == ml.id[k]);
        }
        if (!use) continue;                                   // uniform over the warp (same row)
        const float* hr = h + ((int64_t)k * B + r) * K;
        const float* w = Wl + (int64_t)ml.id[k] * K;
        float v = 0.f;
        for (int j = lane; j < K; j += 32) v = fmaf(hr[j], w[j], v);
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        v += bl[ml.id[k]];
        acc = mode == MB_REDUCE_MIN ? fminf(acc, v) : mode == MB_REDUCE_MAX ? fmaxf(acc, v) : acc + v;
    }
    if (lane) return;
    if (mode == MB_REDUCE_MEAN) acc /= (float)n_sub;
    if (td) acc = reward[r] + (1.f - terminal[r]) * discount * (acc - alpha * log_prob[r]);
    value[r] = acc;
}

// value[b] = reduce over the evaluated members that count for row b (+ optional TD transform)
__global__ void k_redq_reduce(const float* __restrict__ q, int nmem, MemberList ml, int64_t B,
                              const uint8_t* __restrict__ row_members, int n_sub, int mode, int td,
                              const float* __restrict__ reward, const float* __restrict__ terminal,
                              const float* __restrict__ log_prob, float alpha, float discount,
                              float* __restrict__ value) {
    pdl_trigger();
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    float acc = mode == MB_REDUCE_MIN ? INFINITY : mode == MB_REDUCE_MAX ? -INFINITY : 0.f;
    for (int k = 0; k < nmem; ++k) {
        bool use = true;
        if (row_members) {
            use = false;
            for (int sidx = 0; sidx < n_sub; ++sidx) use = use || (row_members[(int64_t)sidx * B + r] == ml.id[k]);
        }
        if (use) {
            const float v = q[(int64_t)k * B + r];
            acc = mode == MB_REDUCE_MIN ? fminf(acc, v) : mode == MB_REDUCE_MAX ? fmaxf(acc, v) : acc + v;
        }
    }
    if (mode == MB_REDUCE_MEAN) acc /= (float)n_sub;
    if (td) acc = reward[r] + (1.f - terminal[r]) * discount * (acc - alpha * log_prob[r]);
    value[r] = acc;
}

// one warp per row: pool E*Q atoms (member-major), bitonic sort, drop the n_drop largest
__global__ void __launch_bounds__(256)
k_tqc_sort(const float* __restrict__ raw /*[E][B][Q]*/, int E, int Q, int64_t B, int n_drop, int npad, int td,
           const float* __restrict__ reward, const float* __restrict__ terminal,
           const float* __restrict__ log_prob, float alpha, float discount, float* __restrict__ value,
           float* __restrict__ atoms) {
    extern __shared__ float pool[];                       // [8 warps][npad]
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= B) return;
    float* row = pool + warp * npad;
    const int NA = E * Q, keep = NA - n_drop;
    for (int t = lane; t < npad; t += 32) {
        float v = INFINITY;
        if (t < NA) { const int e = t / Q, qi = t - e * Q; v = raw[((int64_t)e * B + r) * Q + qi]; }
        row[t] = v;
    }
    __syncwarp();
    if (n_drop > 0) {
        for (int k = 2; k <= npad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < npad / 2; t += 32) {
                    const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
                    const bool up = ((lo & k) == 0);
                    const float a = row[lo], b = row[hi];
                    if ((a > b) == up) { row[lo] = b; row[hi] = a; }
                }
                __syncwarp();
            }
        }
    }
    float sum = 0.f;
    for (int t = lane; t < keep; t += 32) sum += row[t];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0 && value) value[r] = sum / (float)keep;
    if (atoms) {
        float rw = 0.f, sc = 1.f, sh = 0.f;
        if (td) { rw = reward[r]; sc = (1.f - terminal[r]) * discount; sh = alpha * log_prob[r]; }
        for (int t = lane; t < keep; t += 32) atoms[r * keep + t] = td ? rw + sc * (row[t] - sh) : row[t];
    }
}

static int max_width(const mb_mlp_desc& d) {
    int w = 0;
    for (int l = 0; l <= d.num_layers; ++l) w = d.sizes[l] > w ? d.sizes[l] : w;
    return w;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return MB_ECUDA; }
    return MB_OK;
}

template <class Tile>
static int critic_forward_t(const mb_mlp_desc& d, const float* params, const float* obs,
                            const float* act, int O, int A, int64_t B, float* out, cudaStream_t s) {
    int rc = set_smem(k_critic_forward<Tile>, Tile::kSmemBytes);
    if (rc) return rc;
    dim3 grid((unsigned)((B + Tile::TM - 1) / Tile::TM), (unsigned)d.ensemble_size);
    k_critic_forward<Tile><<<grid, kSimtThreads, Tile::kSmemBytes, s>>>(d, params, obs, act, O, A, B, out);
    MB_CUDA_LAUNCH_CHECK("k_critic_forward");
    return MB_OK;
}

int launch_critic_forward(const mb_mlp_desc& d, const float* params, const float* obs,
                          const float* act, int O, int A, int64_t B, float* out, void* ws, cudaStream_t s) {
    const int w = max_width(d);
    if (ws && B * d.ensemble_size <= kLayeredMaxRows) {
        int members[32];
        for (int i = 0; i < d.ensemble_size; ++i) members[i] = i;
        float* res = nullptr;
        return layered_forward(d, params, obs, act, O, A, B, members, d.ensemble_size, (float*)ws, out, &res, s);
    }
    if (w <= SimtTile64::MAXW) return critic_forward_t<SimtTile64>(d, params, obs, act, O, A, B, out, s);
    if (w <= SimtTile32::MAXW) return critic_forward_t<SimtTile32>(d, params, obs, act, O, A, B, out, s);
    set_error("critic width %d > %d unsupported", w, SimtTile32::MAXW);
    return MB_EUNSUPPORTED;
}

template <class Tile>
static int critic_redq_t(const mb_mlp_desc& d, const float* params, const float* obs,
                         const float* act, int O, int A, int64_t B, const MemberList& ml,
                         const uint8_t* row_members, int n_sub, int mode, int td,
                         const float* reward, const float* terminal, const float* log_prob,
                         float alpha, float discount, float* value, cudaStream_t s) {
    const int IN4 = (O + A + 3) & ~3;
    const size_t smem = Tile::kSmemBytes + (size_t)Tile::TM * (IN4 + 1) * sizeof(float);
    int rc = set_smem(k_critic_redq<Tile>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((B + Tile::TM - 1) / Tile::TM);
    k_critic_redq<Tile><<<grid, kSimtThreads, smem, s>>>(d, params, obs, act, O, A, B, ml, row_members,
                                                         n_sub, mode, td, reward, terminal, log_prob,
                                                         alpha, discount, value);
    MB_CUDA_LAUNCH_CHECK("k_critic_redq");
    return MB_OK;
}

int launch_critic_redq(const mb_mlp_desc& d, const float* params, const float* obs, const float* act,
                       int O, int A, int64_t B, const int32_t* members, const uint8_t* row_members,
                       int n_sub, int mode, int td, const float* reward, const float* terminal,
                       const float* log_prob, float alpha, float discount, float* value, void* ws,
                       cudaStream_t s) {
    MemberList ml;
    if (members) {
        ml.n = n_sub;
        for (int i = 0; i < n_sub; ++i) ml.id[i] = members[i];
    } else {
        ml.n = d.ensemble_size;
        for (int i = 0; i < ml.n; ++i) ml.id[i] = i;
        if (!row_members) n_sub = d.ensemble_size;
    }
    const int w = max_width(d);
    if (ws && B * ml.n <= kLayeredMaxRows) {
        float* q = nullptr;
        const int L = d.num_layers;
        if (L >= 2 && d.sizes[L] == 1) {       // Q critics: the 1-wide output layer rides in the reduction
            int rc = layered_forward(d, params, obs, act, O, A, B, ml.id, ml.n, (float*)ws, nullptr, &q, s, L - 1);
            if (rc) return rc;
            launch_pdl(k_redq_reduce_dot, dim3((unsigned)((B + 7) / 8)), dim3(256), 0, s, (const float*)q, d.sizes[L - 1],
                       params + w_offset(d, L - 1), params + b_offset(d, L - 1), ml.n, ml, B, row_members, n_sub, mode,
                       td, reward, terminal, log_prob, alpha, discount, value);
            MB_CUDA_LAUNCH_CHECK("k_redq_reduce_dot");
            return MB_OK;
        }
        int rc = layered_forward(d, params, obs, act, O, A, B, ml.id, ml.n, (float*)ws, nullptr, &q, s);
        if (rc) return rc;
        launch_pdl(k_redq_reduce, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, s, (const float*)q, ml.n, ml, B,
                   row_members, n_sub, mode, td, reward, terminal, log_prob, alpha, discount, value);
        MB_CUDA_LAUNCH_CHECK("k_redq_reduce");
        return MB_OK;
    }
    // small batches are latency bound: 32-row tiles spread a 256-row batch over more SMs
    if (w <= SimtTile64::MAXW && B > 64 * 148)
        return critic_redq_t<SimtTile64>(d, params, obs, act, O, A, B, ml, row_members, n_sub, mode, td,
                                         reward, terminal, log_prob, alpha, discount, value, s);
    if (w <= SimtTile32::MAXW)
        return critic_redq_t<SimtTile32>(d, params, obs, act, O, A, B, ml, row_members, n_sub, mode, td,
                                         reward, terminal, log_prob, alpha, discount, value, s);
    set_error("critic width %d > %d unsupported", w, SimtTile32::MAXW);
    return MB_EUNSUPPORTED;
}

int launch_critic_tqc(const mb_mlp_desc& d, const float* params, const float* obs, const float* act,
                      int O, int A, int64_t B, int n_drop, int td, const float* reward,
                      const float* terminal, const float* log_prob, float alpha, float discount,
                      float* value, float* atoms, void* ws, cudaStream_t s) {
    using Tile = SimtTile32;
    const int w = max_width(d);
    if (ws && B * d.ensemble_size <= kLayeredMaxRows) {
        int members[32];
        for (int i = 0; i < d.ensemble_size; ++i) members[i] = i;
        float* raw = nullptr;
        int rc = layered_forward(d, params, obs, act, O, A, B, members, d.ensemble_size, (float*)ws, nullptr, &raw, s);
        if (rc) return rc;
        const int NAl = d.ensemble_size * d.sizes[d.num_layers];
        int np2 = 32;
        while (np2 < NAl) np2 <<= 1;
        launch_pdl(k_tqc_sort, dim3((unsigned)((B + 7) / 8)), dim3(256), (size_t)8 * np2 * sizeof(float), s,
                   (const float*)raw, d.ensemble_size, d.sizes[d.num_layers], B, n_drop, np2, td, reward, terminal,
                   log_prob, alpha, discount, value, atoms);
        MB_CUDA_LAUNCH_CHECK("k_tqc_sort");
        return MB_OK;
    }
    if (w > Tile::MAXW) { set_error("critic width %d > %d unsupported", w, Tile::MAXW); return MB_EUNSUPPORTED; }
    const int NA = d.ensemble_size * d.sizes[d.num_layers];
    int npad = 32;
    while (npad < NA) npad <<= 1;
    const int IN4 = (O + A + 3) & ~3;
    const size_t smem = Tile::kSmemBytes + (size_t)Tile::TM * (IN4 + npad) * sizeof(float);
    if (smem > 227 * 1024) { set_error("tqc: %d atoms per row do not fit shared memory", NA); return MB_EUNSUPPORTED; }
    int rc = set_smem(k_critic_tqc<Tile>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((B + Tile::TM - 1) / Tile::TM);
    k_critic_tqc<Tile><<<grid, kSimtThreads, smem, s>>>(d, params, obs, act, O, A, B, n_drop, npad, td,
                                                        reward, terminal, log_prob, alpha, discount,
                                                        value, atoms);
    MB_CUDA_LAUNCH_CHECK("k_critic_tqc");
    return MB_OK;
}

}  // namespace mb

// ---- N1: policy action -------------------------------------------------------------------
// GaussianPolicy.action -> MeanLogstdGaussian.forward (mirl/policies/gaussian_policy.py:35-47,
// mirl/torch_modules/gaussian.py:55-97,199-204): out = MLP(obs) [B][2A]; mean, logstd = chunk(out);
// logstd = bound(logstd + extra_bias); std = exp(logstd) * std_scale + min_std; mean *= mean_scale;
// pre = mean + std * eps (eps = the Normal.rsample / sample draw, caller-supplied; NULL = deterministic);
// action = tanh(pre) when squashed; log_prob = sum_j [ Normal.log_prob(pre) - (2 log 2 - softplus(2 pre) -
// softplus(-2 pre)) ].  One thread per (row, action dim); a row's A lanes reduce the log-prob by shuffle.
namespace mb {

constexpr int64_t kPolicyRowsMin = 4096;     // rows from which the one-launch forward (mlp_rows.cu) is used

__global__ void k_policy_head(const float* __restrict__ out, const float* __restrict__ eps, int64_t B, int A,
                              mb_policy_cfg cfg, float* __restrict__ action, float* __restrict__ log_prob,
                              float* __restrict__ mean_out, float* __restrict__ std_out) {
    pdl_trigger();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = t / A;
    const int j = (int)(t - b * A);
    float lp = 0.f;
    if (b < B) {
        float mean = out[b * 2 * A + j] * cfg.mean_scale;
        float ls = out[b * 2 * A + A + j] + cfg.logstd_extra_bias;
        if (cfg.bound_mode == MB_LOGSTD_TANH) ls = (tanhf(ls) + 1.f) * 6.f - 10.f;          // (max - min) / 2 = 6, min = -10
        else if (cfg.bound_mode == MB_LOGSTD_CLAMP) ls = fminf(fmaxf(ls, -10.f), 2.f);
        const float sd = expf(ls) * cfg.std_scale + cfg.min_std;
        const float e = eps ? eps[b * A + j] : 0.f;
        const float pre = mean + sd * e;
        action[b * A + j] = cfg.squashed ? tanhf(pre) : pre;
        if (mean_out) mean_out[b * A + j] = mean;
        if (std_out) std_out[b * A + j] = sd;
        // Normal(mean, sd).log_prob(pre) = -e^2 / 2 - log sd - log sqrt(2 pi)
        lp = -0.5f * e * e - logf(sd) - 0.91893853320467274178f;
        if (cfg.squashed) lp -= 2.f * 0.69314718055994530942f - softplusf_(2.f * pre) - softplusf_(-2.f * pre);
    }
    if (log_prob == nullptr) return;
    // rows do not straddle warps when A divides 32; otherwise fall back to atomics on a zeroed output
    if ((32 % A) == 0) {
        for (int o = 1; o < A; o <<= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
        if (b < B && j == 0) log_prob[b] = lp;
    } else if (b < B) {
        atomicAdd(log_prob + b, lp);
    }
}

int launch_policy_action(const mb_mlp_desc& d, const float* params, const float* obs, int O, int64_t B,
                         const float* eps, const mb_policy_cfg& cfg, float* action, float* log_prob, float* mean,
                         float* std, void* ws, size_t ws_bytes, cudaStream_t s) {
    const int A = d.sizes[d.num_layers] / 2;
    // [out B x 2A | critic workspace]
    float* out = reinterpret_cast<float*>(ws);
    const size_t out_bytes = ((size_t)B * 2 * A * sizeof(float) + 255) & ~size_t(255);
    void* cws = ws_bytes > out_bytes + critic_workspace_bytes(d, B) ? (void*)((char*)ws + out_bytes) : nullptr;
    // rollout batch sizes: the whole network in one launch with the activations resident in shared memory
    // (mlp_rows.cu); small batches keep the per-layer kernels (a handful of 64-row blocks would leave the GPU idle)
    static int rows_on = -1;
    if (rows_on < 0) {
        const char* e = getenv("MB_POLICY_ROWS");
        rows_on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    int rc;
    if (rows_on && B >= kPolicyRowsMin && d.ensemble_size == 1 && mlp_rows_supported(d) &&
        ws_bytes >= out_bytes + mlp_rows_pack_bytes(d)) {
        rc = launch_mlp_rows_forward(d, params, 0, obs, B, out, (char*)ws + out_bytes, s);
    } else {
        rc = launch_critic_forward(d, params, obs, obs, O, 0, B, out, cws, s);
    }
    if (rc) return rc;
    if (log_prob && (32 % A) != 0) cudaMemsetAsync(log_prob, 0, sizeof(float) * B, s);
    const int64_t total = B * A;
    // a warp must hold whole rows for the shuffle reduction: 256 threads = 8 warps, A | 32
    launch_pdl(k_policy_head, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, (const float*)out, eps, B, A, cfg,
               action, log_prob, mean, std);
    MB_CUDA_LAUNCH_CHECK("k_policy_head");
    return MB_OK;
}

}  // namespace mb
