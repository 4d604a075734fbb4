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
                          const float* act, int O, int A, int64_t B, float* out, cudaStream_t s) {
    const int w = max_width(d);
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
                       const float* log_prob, float alpha, float discount, float* value,
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
                      float* value, float* atoms, cudaStream_t s) {
    using Tile = SimtTile32;
    const int w = max_width(d);
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
