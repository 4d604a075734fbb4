// Probabilistic-ensemble kernels, SIMT fp32 family:
//   * row bucketing by drawn member (so a tile runs ONE member instead of E, unlike
//     pe_model.py:163,171-174 which evaluates every member and keeps one),
//   * fused rollout step  (PEModel.step: base_model.py:54-76 + pe_model.py:153-189),
//   * all-member forward / evaluation loss (pe_model.py:110-151, :71-96),
//   * all-elite rollout step with MOPO penalty (pe_model.py:255-282, mopo_collector.py:42-53).
#include "common.cuh"
#include "simt_mlp.cuh"
#include "pe_kernels.h"

namespace mb {

// ------------------------------------------------------------------------------------------
// Bucketing: stable-enough counting sort of row ids by member (E <= 32).
// ------------------------------------------------------------------------------------------
constexpr int kBucketRows = 2048;   // rows per block
constexpr int kMaxE = 32;

__global__ void __launch_bounds__(256)
k_bucket(const uint8_t* __restrict__ idx, int64_t B, int E, int* __restrict__ totals, int* __restrict__ perm) {
    __shared__ int hist[kMaxE], base[kMaxE];
    if (threadIdx.x < kMaxE) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * kBucketRows;
    int e_of[kBucketRows / 256], rank[kBucketRows / 256];
#pragma unroll
    for (int k = 0; k < kBucketRows / 256; ++k) {
        const int64_t r = r0 + k * 256 + threadIdx.x;
        e_of[k] = -1;
        if (r < B) {
            const int e = idx[r];
            e_of[k] = e < E ? e : E - 1;         // out-of-range ids are clamped (memory safety)
            rank[k] = atomicAdd(&hist[e_of[k]], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < kMaxE && hist[threadIdx.x] > 0)
        base[threadIdx.x] = atomicAdd(&totals[threadIdx.x], hist[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBucketRows / 256; ++k)
        if (e_of[k] >= 0) perm[(int64_t)e_of[k] * B + base[e_of[k]] + rank[k]] = (int)(r0 + k * 256 + threadIdx.x);
}

size_t bucket_workspace_bytes(int64_t B, int E) {
    return 256 + (size_t)E * (size_t)B * 4 + 1024;         // totals | one segment per member
}

BucketPlan bucket_rows(const uint8_t* idx, int64_t B, int E, int TM, void* ws, cudaStream_t s) {
    BucketPlan p;
    const int nblk = (int)((B + kBucketRows - 1) / kBucketRows);
    char* w = reinterpret_cast<char*>(ws);
    p.totals = reinterpret_cast<int*>(w);
    p.perm = reinterpret_cast<int*>(w + 256);
    p.seg = B;
    p.max_tiles = (int)((B + TM - 1) / TM + E);
    cudaMemsetAsync(p.totals, 0, kMaxE * sizeof(int), s);
    k_bucket<<<nblk, 256, 0, s>>>(idx, B, E, p.totals, p.perm);
    return p;
}

// ------------------------------------------------------------------------------------------
// Input staging shared by the SIMT kernels: x0[i][0..O+A) = normalised [obs | act] of row i.
// ------------------------------------------------------------------------------------------
template <class Tile>
__device__ __forceinline__ void stage_inputs(const PeDev& m, const float* __restrict__ obs,
                                             const float* __restrict__ act, int cnt,
                                             const int* __restrict__ rows, int64_t row0,
                                             bool normalize, float* x0, float* raw_obs,
                                             int64_t obs_stride = -1, int64_t act_stride = -1,
                                             float* raw_act = nullptr) {
    const int O = m.O, A = m.A, IN = O + A, IN4 = (IN + 3) & ~3;
    const float* nrm = m.norm;
    if (obs_stride < 0) obs_stride = O;
    if (act_stride < 0) act_stride = A;
    for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads) {
        const int i = t / IN4, j = t - i * IN4;
        float v = 0.f;
        if (i < cnt && j < IN) {
            const int64_t r = rows ? (int64_t)rows[i] : row0 + i;
            if (j < O) {
                const float x = obs[r * obs_stride + j];
                if (raw_obs) raw_obs[i * O + j] = x;
                v = normalize ? (x - nrm[j]) / (nrm[O + j] + kNormEps) : x;
            } else {
                const int k = j - O;
                const float x = act[r * act_stride + k];
                if (raw_act) raw_act[i * A + k] = x;
                v = normalize ? (x - nrm[2 * O + k]) / (nrm[2 * O + A + k] + kNormEps) : x;
            }
        }
        x0[i * Tile::AS + j] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Fused rollout step (one member per tile).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSimtThreads, 1)
k_pe_rollout_simt(PeDev m, const float* __restrict__ obs, int64_t obs_stride,
                  const float* __restrict__ act, int64_t act_stride,
                  const float* __restrict__ noise, const int* __restrict__ perm,
                  const int* __restrict__ totals, int64_t seg, int done_kind,
                  float* __restrict__ next_obs, float* __restrict__ reward,
                  float* __restrict__ done, RowsOut rows_out) {
    using Tile = SimtTile64;
    __shared__ TileIndex tix;
    if (threadIdx.x < 32) tile_index_build(tix, totals, m.mlp.ensemble_size, Tile::TM);
    __syncthreads();
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    float* raw = wbuf + 2 * Tile::kWFloats;          // [TM][O] raw observations
    float* nob = raw + Tile::TM * m.O;               // [TM][O] next observations
    float* rawa = nob + Tile::TM * m.O;              // [TM][A] raw actions (row output)
    const int O = m.O, D = O + 1, A = m.A;
    const int64_t W = rows_out.stride;               // destination row stride (floats)
    const int ntiles = tix.toff[kMaxMembers];
    const float* nrm = m.norm;
    const float* dmean = nrm + 2 * O + 2 * m.A;
    const float* dstd = dmean + O;
    const float* mxp = m.params + logstd_offset(m.mlp, 0);
    const float* mnp = m.params + logstd_offset(m.mlp, 1);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int4 ti = tile_at(tix, Tile::TM, tile);
        const int e = ti.x, cnt = ti.z;
        const int* rows = perm + (size_t)e * seg + ti.y;
        stage_inputs<Tile>(m, obs, act, cnt, rows, 0, true, buf0, raw, obs_stride, act_stride, rawa);
        __syncthreads();
        const float* res = Tile::forward(m.mlp, m.params, e, buf0, buf1, wbuf, nullptr, cnt);
        for (int t = threadIdx.x; t < cnt * D; t += kSimtThreads) {
            const int i = t / D, j = t - i * D;
            const int64_t r = rows[i];
            // destination row: batch order, or bucket order (position of the row in its tile)
            const int64_t dr = rows_out.row_offset + (rows_out.ordered ? r : (int64_t)tix.roff[e] + ti.y + i);
            const float mu = res[i * Tile::AS + j];
            const float ls = clamp_logstd(res[i * Tile::AS + D + j], mxp[e * D + j], mnp[e * D + j]);
            const float s = noise[r * D + j] * expf(ls) + mu;
            if (j < O) {
                const float delta = s * (dstd[j] + kNormEps) + dmean[j];
                const float v = raw[i * O + j] + delta;
                nob[i * O + j] = v;
                if (next_obs) next_obs[r * O + j] = v;
#pragma unroll
                for (int k = 0; k < kMaxRowDest; ++k) {
                    if (k >= rows_out.n) continue;
                    float* dst = rows_out.dst[k] + (size_t)dr * W;
                    dst[j] = raw[i * O + j];
                    dst[O + A + j] = v;
                }
            } else {
                if (reward) reward[r] = s;
#pragma unroll
                for (int k = 0; k < kMaxRowDest; ++k)
                    if (k < rows_out.n) rows_out.dst[k][(size_t)dr * W + 2 * O + A] = s;
            }
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int64_t r = rows[threadIdx.x];
            const int64_t dr = rows_out.row_offset +
                               (rows_out.ordered ? r : (int64_t)tix.roff[e] + ti.y + (int64_t)threadIdx.x);
            const float dn = done_fn(done_kind, nob + threadIdx.x * O, O);
            if (done) done[r] = dn;
#pragma unroll
            for (int k = 0; k < kMaxRowDest; ++k) {
                if (k >= rows_out.n) continue;
                float* dst = rows_out.dst[k] + (size_t)dr * W;
                dst[2 * O + A + 1] = dn;
                for (int j = 0; j < A; ++j) dst[O + j] = rawa[threadIdx.x * A + j];
                for (int j = 2 * O + A + 2; j < W; ++j) dst[j] = 0.f;          // row padding
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// All-member forward; optional mean/logstd outputs and optional Gaussian-NLL accumulation.
// grid = (row tiles, E).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSimtThreads, 1)
k_pe_forward_simt(PeDev m, const float* __restrict__ obs, const float* __restrict__ act,
                  int per_member, int normalize, int64_t B, float* __restrict__ mean_out,
                  float* __restrict__ logstd_out, const float* __restrict__ delta,
                  const float* __restrict__ reward, const float* __restrict__ done_in,
                  int ignore_terminal, float reward_coef, float* __restrict__ ensemble_loss) {
    using Tile = SimtTile64;
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    __shared__ float red[kSimtThreads / 32];
    const int O = m.O, A = m.A, D = O + 1;
    const int e = blockIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * Tile::TM;
    const int cnt = (int)min((int64_t)Tile::TM, B - row0);
    const int64_t moff = per_member ? (int64_t)e * B : 0;
    const float* nrm = m.norm;
    const float* dmean = nrm + 2 * O + 2 * A;
    const float* dstd = dmean + O;
    const float* mxp = m.params + logstd_offset(m.mlp, 0);
    const float* mnp = m.params + logstd_offset(m.mlp, 1);

    stage_inputs<Tile>(m, obs + moff * O, act + moff * A, cnt, nullptr, row0, normalize != 0, buf0,
                       nullptr);
    __syncthreads();
    const float* res = Tile::forward(m.mlp, m.params, e, buf0, buf1, wbuf, nullptr, cnt);
    float part = 0.f;
    for (int t = threadIdx.x; t < cnt * D; t += kSimtThreads) {
        const int i = t / D, j = t - i * D;
        const int64_t r = row0 + i;
        const float mu = res[i * Tile::AS + j];
        const float ls = clamp_logstd(res[i * Tile::AS + D + j], mxp[e * D + j], mnp[e * D + j]);
        if (mean_out) {
            mean_out[((int64_t)e * B + r) * D + j] = mu;
            logstd_out[((int64_t)e * B + r) * D + j] = ls;
        }
        if (ensemble_loss) {
            const int64_t rb = moff + r;
            float tgt, coef;
            if (j < O) {
                const float dl = delta[rb * O + j];
                tgt = normalize ? (dl - dmean[j]) / (dstd[j] + kNormEps) : dl;
                coef = 1.f;
            } else {
                tgt = reward[rb];
                coef = reward_coef;
            }
            const float diff = tgt - mu;
            float nll = 0.5f * (kLog2Pi + 2.f * ls + diff * diff * expf(-2.f * ls)) * coef;
            if (ignore_terminal) nll *= (1.f - done_in[rb]);
            part += nll;
        }
    }
    if (ensemble_loss) {
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < kSimtThreads / 32; ++w) s += red[w];
            atomicAdd(ensemble_loss + e, s / (float)B);
        }
    }
}

// ------------------------------------------------------------------------------------------
// All-elite rollout step: sampled transition + per-elite distribution info + MOPO penalty.
// ------------------------------------------------------------------------------------------
struct EliteList { int n; int id[kMaxE]; };

__global__ void __launch_bounds__(kSimtThreads, 1)
k_pe_rollout_dist_simt(PeDev m, EliteList el, const float* __restrict__ obs,
                       const float* __restrict__ act, const uint8_t* __restrict__ member_idx,
                       const float* __restrict__ noise, int64_t B, int done_kind,
                       float* __restrict__ next_obs, float* __restrict__ reward,
                       float* __restrict__ done, float* __restrict__ delta_mean,
                       float* __restrict__ delta_std, float* __restrict__ next_obs_mean,
                       float* __restrict__ reward_mean, float* __restrict__ reward_std,
                       int penalty_kind, float penalty_coef, float* __restrict__ penalty,
                       float* __restrict__ penalized_reward) {
    using Tile = SimtTile64;
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    const int O = m.O, A = m.A, D = O + 1, TM = Tile::TM;
    float* raw = wbuf + 2 * Tile::kWFloats;   // [TM][O]
    float* x0 = raw + TM * O;                 // [TM][IN4] normalised inputs (kept across elites)
    const int IN4 = (O + A + 3) & ~3;
    float* nob = x0 + TM * IN4;               // [TM][O] sampled next obs
    float* wmean = nob + TM * O;              // Welford mean of next_obs_mean over elites
    float* wm2 = wmean + TM * O;              // Welford M2
    float* svar = wm2 + TM * O;               // sum over elites of std^2
    float* sq = svar + TM * O;                // this elite's std^2 per (row, j)
    float* mxn = sq + TM * O;                 // [TM] max over elites of ||std||
    float* rew = mxn + TM;                    // [TM] sampled reward
    const int64_t row0 = (int64_t)blockIdx.x * TM;
    const int cnt = (int)min((int64_t)TM, B - row0);
    const float* nrm = m.norm;
    const float* dmean = nrm + 2 * O + 2 * A;
    const float* dstd = dmean + O;
    const float* mxp = m.params + logstd_offset(m.mlp, 0);
    const float* mnp = m.params + logstd_offset(m.mlp, 1);

    stage_inputs<Tile>(m, obs, act, cnt, nullptr, row0, true, buf0, raw);
    for (int t = threadIdx.x; t < TM * O; t += kSimtThreads) { wmean[t] = 0.f; wm2[t] = 0.f; svar[t] = 0.f; }
    if (threadIdx.x < TM) mxn[threadIdx.x] = 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < TM * IN4; t += kSimtThreads)
        x0[t] = buf0[(t / IN4) * Tile::AS + (t % IN4)];
    __syncthreads();

    for (int k = 0; k < el.n; ++k) {
        const int e = el.id[k];
        if (k > 0) {
            for (int t = threadIdx.x; t < TM * IN4; t += kSimtThreads)
                buf0[(t / IN4) * Tile::AS + (t % IN4)] = x0[t];
            __syncthreads();
        }
        const float* res = Tile::forward(m.mlp, m.params, e, buf0, buf1, wbuf, nullptr, cnt);
        for (int t = threadIdx.x; t < cnt * D; t += kSimtThreads) {
            const int i = t / D, j = t - i * D;
            const int64_t r = row0 + i;
            const float mu = res[i * Tile::AS + j];
            const float ls = clamp_logstd(res[i * Tile::AS + D + j], mxp[e * D + j], mnp[e * D + j]);
            const float sd = expf(ls);
            const bool mine = member_idx[r] == (uint8_t)e;
            const float s = mine ? noise[r * D + j] * sd + mu : 0.f;
            const int64_t o3 = ((int64_t)k * B + r);
            if (j < O) {
                const float scale = dstd[j] + kNormEps;
                const float dm = mu * scale + dmean[j];
                const float ds = sd * scale;
                const float nm = dm + raw[i * O + j];
                if (delta_mean) delta_mean[o3 * O + j] = dm;
                if (delta_std) delta_std[o3 * O + j] = ds;
                if (next_obs_mean) next_obs_mean[o3 * O + j] = nm;
                // Welford update over elites (torch.var, unbiased) + mean of variances
                const float dlt = nm - wmean[i * O + j];
                const float nmean = wmean[i * O + j] + dlt / (float)(k + 1);
                wm2[i * O + j] += dlt * (nm - nmean);
                wmean[i * O + j] = nmean;
                svar[i * O + j] += ds * ds;
                sq[i * O + j] = ds * ds;
                if (mine) {
                    const float v = raw[i * O + j] + (s * scale + dmean[j]);
                    nob[i * O + j] = v;
                    next_obs[r * O + j] = v;
                }
            } else {
                if (reward_mean) reward_mean[o3] = mu;
                if (reward_std) reward_std[o3] = sd;
                if (mine) { rew[i] = s; reward[r] = s; }
            }
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            float n2 = 0.f;
            for (int j = 0; j < O; ++j) n2 += sq[threadIdx.x * O + j];
            mxn[threadIdx.x] = fmaxf(mxn[threadIdx.x], sqrtf(n2));
        }
        __syncthreads();
    }
    if (threadIdx.x < cnt) {
        const int i = threadIdx.x;
        const int64_t r = row0 + i;
        done[r] = done_fn(done_kind, nob + i * O, O);
        if (penalty_kind != MB_PENALTY_NONE) {
            float pen;
            if (penalty_kind == MB_PENALTY_MAX_VAR) {
                pen = mxn[i];
            } else {
                pen = 0.f;
                for (int j = 0; j < O; ++j)
                    pen += svar[i * O + j] / (float)el.n + wm2[i * O + j] / (float)(el.n - 1);
            }
            if (penalty) penalty[r] = pen;
            if (penalized_reward) penalized_reward[r] = rew[i] - penalty_coef * pen;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------
static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return MB_ECUDA; }
    return MB_OK;
}

int launch_pe_rollout_simt(const PeDev& m, const float* obs, int64_t obs_stride, const float* act,
                           int64_t act_stride, const uint8_t* idx, const float* noise, int64_t B,
                           int done_kind, float* next_obs, float* reward, float* done,
                           const RowsOut& rows, void* ws, cudaStream_t s) {
    using Tile = SimtTile64;
    BucketPlan p = bucket_rows(idx, B, m.mlp.ensemble_size, Tile::TM, ws, s);
    MB_CUDA_LAUNCH_CHECK("bucket_rows");
    const size_t smem = Tile::kSmemBytes + (size_t)Tile::TM * (2 * m.O + m.A) * sizeof(float);
    int rc = set_smem(k_pe_rollout_simt, smem);
    if (rc) return rc;
    const int grid = (int)min((int64_t)p.max_tiles, (int64_t)sm_count() * 4);
    prof_begin(s);
    k_pe_rollout_simt<<<grid, kSimtThreads, smem, s>>>(m, obs, obs_stride, act, act_stride, noise, p.perm,
                                                       p.totals, p.seg, done_kind, next_obs, reward, done,
                                                       rows);
    prof_end(s);
    MB_CUDA_LAUNCH_CHECK("k_pe_rollout_simt");
    return MB_OK;
}

int launch_pe_forward_simt(const PeDev& m, const float* obs, const float* act, int per_member,
                           int normalize, int64_t B, float* mean, float* logstd, const float* delta,
                           const float* reward, const float* done, int ignore_terminal,
                           float reward_coef, float* ensemble_loss, cudaStream_t s) {
    using Tile = SimtTile64;
    int rc = set_smem(k_pe_forward_simt, Tile::kSmemBytes);
    if (rc) return rc;
    dim3 grid((unsigned)((B + Tile::TM - 1) / Tile::TM), (unsigned)m.mlp.ensemble_size);
    k_pe_forward_simt<<<grid, kSimtThreads, Tile::kSmemBytes, s>>>(
        m, obs, act, per_member, normalize, B, mean, logstd, delta, reward, done, ignore_terminal,
        reward_coef, ensemble_loss);
    MB_CUDA_LAUNCH_CHECK("k_pe_forward_simt");
    return MB_OK;
}

// ---- MOPO / return_distribution tail over a precomputed all-elite forward --------------------
// fwd[k][b][2D] = [mean | clamped log-std] of elite k for row b (normalised space, written by the
// tcgen05 forward).  One warp per row, lanes = output columns (D <= 32): the same outputs as
// k_pe_rollout_dist_simt (pe_model.py:255-282, mopo_collector.py:42-53).
__global__ void __launch_bounds__(256)
k_pe_dist_finish(PeDev m, EliteList el, const float* __restrict__ fwd, const float* __restrict__ obs,
                 const uint8_t* __restrict__ member_idx, const float* __restrict__ noise, int64_t B,
                 int done_kind, float* __restrict__ next_obs, float* __restrict__ reward,
                 float* __restrict__ done, float* __restrict__ delta_mean, float* __restrict__ delta_std,
                 float* __restrict__ next_obs_mean, float* __restrict__ reward_mean,
                 float* __restrict__ reward_std, int penalty_kind, float penalty_coef,
                 float* __restrict__ penalty, float* __restrict__ penalized_reward) {
    __shared__ float nrow[8][32];
    const int O = m.O, A = m.A, D = O + 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* dmean = m.norm + 2 * O + 2 * A;
    const float* dstd = dmean + O;
    const int j = lane;
    const bool in_o = j < O, in_d = j < D;
    const float scale = in_o ? dstd[j] + kNormEps : 0.f;
    const float dm0 = in_o ? dmean[j] : 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 8 + w; r < B; r += (int64_t)gridDim.x * 8) {
        const float raw = in_o ? obs[r * O + j] : 0.f;
        const int mine_e = member_idx[r];
        const float eps = in_d ? noise[r * D + j] : 0.f;
        float wmean = 0.f, wm2 = 0.f, svar = 0.f, mxn = 0.f, rew = 0.f;
        nrow[w][lane] = 0.f;
        for (int k = 0; k < el.n; ++k) {
            const float* f = fwd + ((int64_t)k * B + r) * (2 * D);
            const float mu = in_d ? f[j] : 0.f;
            const float sd = in_d ? expf(f[D + j]) : 0.f;
            const bool mine = el.id[k] == mine_e;
            const float sv = eps * sd + mu;
            const int64_t o3 = (int64_t)k * B + r;
            float sq = 0.f;
            if (in_o) {
                const float dmv = mu * scale + dm0;
                const float ds = sd * scale;
                const float nm = dmv + raw;
                if (delta_mean) delta_mean[o3 * O + j] = dmv;
                if (delta_std) delta_std[o3 * O + j] = ds;
                if (next_obs_mean) next_obs_mean[o3 * O + j] = nm;
                const float dlt = nm - wmean;                      // Welford over elites (torch.var, unbiased)
                const float nmean = wmean + dlt / (float)(k + 1);
                wm2 += dlt * (nm - nmean);
                wmean = nmean;
                sq = ds * ds;
                svar += sq;
                if (mine) {
                    const float v = raw + (sv * scale + dm0);
                    nrow[w][j] = v;
                    next_obs[r * O + j] = v;
                }
            } else if (j == O) {
                if (reward_mean) reward_mean[o3] = mu;
                if (reward_std) reward_std[o3] = sd;
                if (mine) { rew = sv; reward[r] = sv; }
            }
            for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            mxn = fmaxf(mxn, sqrtf(sq));
        }
        __syncwarp();
        float pen = in_o ? svar / (float)el.n + wm2 / (float)(el.n - 1) : 0.f;
        for (int o = 16; o; o >>= 1) pen += __shfl_xor_sync(0xffffffffu, pen, o);
        rew = __shfl_sync(0xffffffffu, rew, O);
        if (lane == 0) {
            done[r] = done_fn(done_kind, nrow[w], O);
            if (penalty_kind != MB_PENALTY_NONE) {
                const float p = penalty_kind == MB_PENALTY_MAX_VAR ? mxn : pen;
                if (penalty) penalty[r] = p;
                if (penalized_reward) penalized_reward[r] = rew - penalty_coef * p;
            }
        }
        __syncwarp();
    }
}

int launch_pe_dist_finish(const PeDev& m, const int32_t* elites, int n_elite, const float* fwd, const float* obs,
                          const uint8_t* idx, const float* noise, int64_t B, int done_kind, float* next_obs,
                          float* reward, float* done, float* delta_mean, float* delta_std, float* next_obs_mean,
                          float* reward_mean, float* reward_std, int penalty_kind, float penalty_coef,
                          float* penalty, float* penalized_reward, cudaStream_t s) {
    EliteList el;
    el.n = n_elite;
    for (int i = 0; i < n_elite; ++i) el.id[i] = elites[i];
    const int64_t blocks = (B + 7) / 8;
    const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
    k_pe_dist_finish<<<grid, 256, 0, s>>>(m, el, fwd, obs, idx, noise, B, done_kind, next_obs, reward, done,
                                          delta_mean, delta_std, next_obs_mean, reward_mean, reward_std,
                                          penalty_kind, penalty_coef, penalty, penalized_reward);
    MB_CUDA_LAUNCH_CHECK("k_pe_dist_finish");
    return MB_OK;
}

int launch_pe_rollout_dist_simt(const PeDev& m, const int32_t* elites, int n_elite, const float* obs,
                                const float* act, const uint8_t* idx, const float* noise, int64_t B,
                                int done_kind, float* next_obs, float* reward, float* done,
                                float* delta_mean, float* delta_std, float* next_obs_mean,
                                float* reward_mean, float* reward_std, int penalty_kind,
                                float penalty_coef, float* penalty, float* penalized_reward,
                                cudaStream_t s) {
    using Tile = SimtTile64;
    EliteList el;
    el.n = n_elite;
    for (int i = 0; i < n_elite; ++i) el.id[i] = elites[i];
    const int O = m.O, IN4 = (m.O + m.A + 3) & ~3;
    const size_t extra = (size_t)Tile::TM * (6 * O + IN4 + 2) * sizeof(float);
    const size_t smem = Tile::kSmemBytes + extra;
    int rc = set_smem(k_pe_rollout_dist_simt, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((B + Tile::TM - 1) / Tile::TM);
    k_pe_rollout_dist_simt<<<grid, kSimtThreads, smem, s>>>(
        m, el, obs, act, idx, noise, B, done_kind, next_obs, reward, done, delta_mean, delta_std,
        next_obs_mean, reward_mean, reward_std, penalty_kind, penalty_coef, penalty, penalized_reward);
    MB_CUDA_LAUNCH_CHECK("k_pe_rollout_dist_simt");
    return MB_OK;
}

}  // namespace mb
