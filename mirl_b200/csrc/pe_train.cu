// Model-training step of the probabilistic ensemble, SIMT fp32 family.
//
// One PEModelTrainer.train_from_torch_batch (mirl/agents/mbpo/pe_model_trainer.py:85-106):
//   forward  (PEModel.compute_loss, pe_model.py:285-307 -> :71-143)  -> k_train_fwd
//   loss.backward()  (autograd over ~60 ATen ops, 9 GEMMs)            -> k_train_gemm<DW|DX>
//   torch.optim.Adam.step() over 84 tensors (pe_model_trainer.py:47-51,99) -> k_adam
// Gradient formulae: SURVEY.md Appendix A.3 (validated against autograd, see
// tests/test_oracle_golden.py::test_analytic_gradients_match_autograd_of_reference).
// Members are independent (the loss is a plain sum over members, pe_model.py:95-96), so a call
// may cover any member slice [e0, e1) -- that is the multi-GPU partition; no collective.
#include "common.cuh"
#include "simt_mlp.cuh"
#include "layer_gemm.cuh"
#include "pe_kernels.h"
#include "tc_gemm.h"
#include <stdlib.h>

namespace mb {

struct TrainWs {
    float* x0;                     // [E][B][IN] normalised inputs
    float* z[MB_MAX_LAYERS];       // [E][B][sizes[l+1]] pre-activations, l < L-1
    float* gA;                     // [E][B][maxW] dL/dz ping
    float* gB;                     // [E][B][maxW] dL/dz pong
    float* gl[MB_MAX_LAYERS];      // small-batch path: dL/d(out of layer l), kept for the batched dW launch
    float* grads;                  // packed blob
};

static int max_width(const mb_mlp_desc& d) {
    int w = 0;
    for (int l = 0; l <= d.num_layers; ++l) w = d.sizes[l] > w ? d.sizes[l] : w;
    return w;
}

// Small batches keep every layer's delta so that all weight-gradient GEMMs go out as ONE launch.
constexpr int64_t kSmallBatchRows = 65536;
static bool kept_deltas(const mb_mlp_desc& d, int64_t B) { return B * d.ensemble_size <= kSmallBatchRows; }

size_t train_workspace_bytes(const mb_mlp_desc& d, int64_t B) {
    const int64_t E = d.ensemble_size;
    int64_t f = round_up4(E * B * d.sizes[0]);
    for (int l = 0; l + 1 < d.num_layers; ++l) f += round_up4(E * B * d.sizes[l + 1]);
    f += (kept_deltas(d, B) ? d.num_layers + 1 : 2) * round_up4(E * B * max_width(d));
    f += param_count(d, true);
    return (size_t)f * sizeof(float) + 256;
}

static TrainWs carve(const mb_mlp_desc& d, int64_t B, void* ws) {
    TrainWs w;
    const int64_t E = d.ensemble_size;
    float* p = reinterpret_cast<float*>(ws);
    w.x0 = p; p += round_up4(E * B * d.sizes[0]);
    for (int l = 0; l < MB_MAX_LAYERS; ++l) w.z[l] = nullptr;
    for (int l = 0; l + 1 < d.num_layers; ++l) { w.z[l] = p; p += round_up4(E * B * d.sizes[l + 1]); }
    w.gA = p; p += round_up4(E * B * max_width(d));
    w.gB = p; p += round_up4(E * B * max_width(d));
    for (int l = 0; l < MB_MAX_LAYERS; ++l) w.gl[l] = nullptr;
    if (kept_deltas(d, B)) {
        w.gl[d.num_layers - 1] = w.gA;                       // head delta
        for (int l = 0; l + 1 < d.num_layers; ++l) { w.gl[l] = p; p += round_up4(E * B * max_width(d)); }
    }
    w.grads = p;
    return w;
}

struct ZPtrs { float* p[MB_MAX_LAYERS]; };

// ---- forward + loss + dL/d(out) ------------------------------------------------------------
// grid = (row tiles, members of the slice); batch tensors are [E][B][.]
__global__ void __launch_bounds__(kSimtThreads, 1)
k_train_fwd(PeDev m, mb_train_cfg cfg, int e0, const float* __restrict__ obs,
            const float* __restrict__ act, const float* __restrict__ delta,
            const float* __restrict__ reward, const float* __restrict__ done_in, int64_t B,
            float* __restrict__ x0s, ZPtrs zs, float* __restrict__ gout, float* __restrict__ grads,
            float* __restrict__ ensemble_loss, float* __restrict__ loss_terms) {
    using Tile = SimtTile64;
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + Tile::kActFloats;
    float* wbuf = buf1 + Tile::kActFloats;
    __shared__ float red[kSimtThreads / 32];
    __shared__ float gmx[64], gmn[64];
    const int O = m.O, A = m.A, D = O + 1, IN = O + A, IN4 = (IN + 3) & ~3;
    const int L = m.mlp.num_layers;
    const int e = e0 + blockIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * Tile::TM;
    const int cnt = (int)min((int64_t)Tile::TM, B - row0);
    const int64_t mrow = (int64_t)e * B;
    const float* nrm = m.norm;
    const float* dmean = nrm + 2 * O + 2 * A;
    const float* dstd = dmean + O;
    const float* mxp = m.params + logstd_offset(m.mlp, 0) + e * D;
    const float* mnp = m.params + logstd_offset(m.mlp, 1) + e * D;

    // stage normalised inputs (Normalizer.process, normalizer.py:17-18) and keep a copy for dW_0
    for (int t = threadIdx.x; t < Tile::TM * IN4; t += kSimtThreads) {
        const int i = t / IN4, j = t - i * IN4;
        float v = 0.f;
        if (i < cnt && j < IN) {
            const int64_t r = mrow + row0 + i;
            if (j < O) v = (obs[r * O + j] - nrm[j]) / (nrm[O + j] + kNormEps);
            else v = (act[r * A + (j - O)] - nrm[2 * O + (j - O)]) / (nrm[2 * O + A + (j - O)] + kNormEps);
            x0s[r * IN + j] = v;
        }
        buf0[i * Tile::AS + j] = v;
    }
    if (threadIdx.x < 64) { gmx[threadIdx.x] = 0.f; gmn[threadIdx.x] = 0.f; }
    __syncthreads();

    float* zsave[MB_MAX_LAYERS];
    for (int l = 0; l < L; ++l)
        zsave[l] = (l + 1 < L) ? zs.p[l] + (mrow + row0) * m.mlp.sizes[l + 1] : nullptr;
    const float* res = Tile::forward(m.mlp, m.params, e, buf0, buf1, wbuf, zsave, cnt);

    const float invB = 1.f / (float)B;
    float part = 0.f;
    for (int t = threadIdx.x; t < cnt * D; t += kSimtThreads) {
        const int i = t / D, j = t - i * D;
        const int64_t r = mrow + row0 + i;
        const float mu = res[i * Tile::AS + j], raw = res[i * Tile::AS + D + j];
        const float mx = mxp[j], mn = mnp[j];
        const float a = mx - raw;
        const float ls1 = mx - softplusf_(a);
        const float c = ls1 - mn;
        const float ls = mn + softplusf_(c);
        float tgt, coef;
        if (j < O) { tgt = (delta[r * O + j] - dmean[j]) / (dstd[j] + kNormEps); coef = 1.f; }
        else { tgt = reward[r]; coef = cfg.reward_coef; }
        const float mask = cfg.ignore_terminal_state ? 1.f - done_in[r] : 1.f;
        const float inv_var = expf(-2.f * ls);
        const float diff = tgt - mu;
        const float w = mask * coef * invB;
        part += w * 0.5f * (kLog2Pi + 2.f * ls + diff * diff * inv_var);
        const float g_ls = w * (1.f - diff * diff * inv_var);
        const float sa = softplus_gradf_(a), sb = softplus_gradf_(c);
        gout[r * (2 * D) + j] = -w * diff * inv_var;
        gout[r * (2 * D) + D + j] = g_ls * sb * sa;
        atomicAdd(&gmx[j], g_ls * sb * (1.f - sa));
        atomicAdd(&gmn[j], g_ls * (1.f - sb));
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kSimtThreads / 32; ++w) s += red[w];
        atomicAdd(ensemble_loss + e, s);
        atomicAdd(loss_terms + 0, s);
    }
    if (threadIdx.x < D) {
        const int j = threadIdx.x;
        float gx = gmx[j], gn = gmn[j];
        if (blockIdx.x == 0) {   // bound regulariser, once per member (pe_model.py:81,105-108)
            gx += cfg.bound_reg_coef;
            gn -= cfg.bound_reg_coef;
            atomicAdd(loss_terms + 2, cfg.bound_reg_coef * (mxp[j] - mnp[j]));
        }
        atomicAdd(grads + logstd_offset(m.mlp, 0) + e * D + j, gx);
        atomicAdd(grads + logstd_offset(m.mlp, 1) + e * D + j, gn);
    }
}

// ---- small-batch forward path: prep + one launch per layer (layer_gemm.cuh) + loss head -------
// x0[e][b][:] = normalised [obs | act] (Normalizer.process, normalizer.py:17-18); also zeroes the
// atomically accumulated outputs of the member slice so the step needs no memset nodes.
__global__ void k_train_prep(PeDev m, int e0, int ne, const float* __restrict__ obs,
                             const float* __restrict__ act, int64_t B, float* __restrict__ x0s,
                             float* __restrict__ grads, float* __restrict__ ensemble_loss,
                             float* __restrict__ loss_terms) {
    pdl_trigger();
    pdl_wait();
    const int O = m.O, A = m.A, IN = O + A, D = O + 1;
    const float* nrm = m.norm;
    const int64_t total = (int64_t)ne * B * IN;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t = gid; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t rowl = t / IN;
        const int j = (int)(t - rowl * IN);
        const int64_t r = (int64_t)e0 * B + rowl;
        float v;
        if (j < O) v = (obs[r * O + j] - nrm[j]) / (nrm[O + j] + kNormEps);
        else v = (act[r * A + (j - O)] - nrm[2 * O + (j - O)]) / (nrm[2 * O + A + (j - O)] + kNormEps);
        x0s[r * IN + j] = v;
    }
    if (gid < ne) ensemble_loss[e0 + gid] = 0.f;
    if (gid < 3) loss_terms[gid] = 0.f;
    // grid-stride: the grid is sized from ne*B*IN, which is smaller than 2*ne*D for B = 1 with a wide D
    for (int64_t t = gid; t < 2 * (int64_t)ne * D; t += (int64_t)gridDim.x * blockDim.x) {
        const int which = (int)(t / (ne * D)), k = (int)(t % (ne * D));
        grads[logstd_offset(m.mlp, which) + (int64_t)e0 * D + k] = 0.f;
    }
}

// Gaussian NLL head on the last layer's raw output out[e][b][2D]: loss, dL/d(out), log-std bound
// gradients.  grid = (row chunks of kLossRows, members of the slice).
constexpr int kLossRows = 32;       // rows per CTA: spreads a 256-row batch over 8 x members CTAs
__global__ void __launch_bounds__(256)
k_train_loss(PeDev m, mb_train_cfg cfg, int e0, const float* __restrict__ out,
             const float* __restrict__ delta, const float* __restrict__ reward,
             const float* __restrict__ done_in, int64_t B, float* __restrict__ gout,
             float* __restrict__ grads, float* __restrict__ ensemble_loss, float* __restrict__ loss_terms) {
    __shared__ float red[8];
    __shared__ float gmx[64], gmn[64];
    pdl_trigger();
    pdl_wait();
    const int O = m.O, A = m.A, D = O + 1;
    const int e = e0 + blockIdx.y;
    const int64_t mrow = (int64_t)e * B;
    const int64_t row0 = (int64_t)blockIdx.x * kLossRows;
    const int cnt = (int)min((int64_t)kLossRows, B - row0);
    const float* nrm = m.norm;
    const float* dmean = nrm + 2 * O + 2 * A;
    const float* dstd = dmean + O;
    const float* mxp = m.params + logstd_offset(m.mlp, 0) + e * D;
    const float* mnp = m.params + logstd_offset(m.mlp, 1) + e * D;
    if (threadIdx.x < 64) { gmx[threadIdx.x] = 0.f; gmn[threadIdx.x] = 0.f; }
    __syncthreads();
    const float invB = 1.f / (float)B;
    float part = 0.f;
    for (int t = threadIdx.x; t < cnt * D; t += 256) {
        const int i = t / D, j = t - i * D;
        const int64_t r = mrow + row0 + i;
        const float mu = out[r * (2 * D) + j], raw = out[r * (2 * D) + D + j];
        const float mx = mxp[j], mn = mnp[j];
        const float a = mx - raw;
        const float ls1 = mx - softplusf_(a);
        const float c = ls1 - mn;
        const float ls = mn + softplusf_(c);
        float tgt, coef;
        if (j < O) { tgt = (delta[r * O + j] - dmean[j]) / (dstd[j] + kNormEps); coef = 1.f; }
        else { tgt = reward[r]; coef = cfg.reward_coef; }
        const float mask = cfg.ignore_terminal_state ? 1.f - done_in[r] : 1.f;
        const float inv_var = expf(-2.f * ls);
        const float diff = tgt - mu;
        const float w = mask * coef * invB;
        part += w * 0.5f * (kLog2Pi + 2.f * ls + diff * diff * inv_var);
        const float g_ls = w * (1.f - diff * diff * inv_var);
        const float sa = softplus_gradf_(a), sb = softplus_gradf_(c);
        gout[r * (2 * D) + j] = -w * diff * inv_var;
        gout[r * (2 * D) + D + j] = g_ls * sb * sa;
        atomicAdd(&gmx[j], g_ls * sb * (1.f - sa));
        atomicAdd(&gmn[j], g_ls * (1.f - sb));
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
        for (int w = 0; w < 8; ++w) sum += red[w];
        atomicAdd(ensemble_loss + e, sum);
        atomicAdd(loss_terms + 0, sum);
    }
    if (threadIdx.x < D) {
        const int j = threadIdx.x;
        float gx = gmx[j], gn = gmn[j];
        if (blockIdx.x == 0) {   // bound regulariser, once per member (pe_model.py:81,105-108)
            gx += cfg.bound_reg_coef;
            gn -= cfg.bound_reg_coef;
            atomicAdd(loss_terms + 2, cfg.bound_reg_coef * (mxp[j] - mnp[j]));
        }
        atomicAdd(grads + logstd_offset(m.mlp, 0) + e * D + j, gx);
        atomicAdd(grads + logstd_offset(m.mlp, 1) + e * D + j, gn);
    }
}

// ---- backward GEMMs ------------------------------------------------------------------------
// DW: dW[k][n] = sum_b h[b][k] * g[b][n] + wd * W[k][n];  db[n] = sum_b g[b][n]
// DX: gprev[b][k] = (sum_n g[b][n] * W[k][n]) * swish'(z[b][k])
// Same structure as layer_gemm.cuh: 64x64 output tile, reduction in chunks of 32 with the next
// chunk's operands prefetched into registers (the grids are ~1 CTA per SM).
constexpr int kGT = 64;    // output tile edge
constexpr int kGR = 32;    // reduction chunk
constexpr int kGLD = kGT + 4;
constexpr int kGPer = kGT * kGR / 256;   // operand elements per thread per chunk

struct GemmLayer {
    int K, N;                      // layer fan-in / fan-out
    const float* hin;              // DW: x0 or z_{l-1}  [E][B][K]
    int hin_is_z;
    const float* g;                // delta of this layer's output [E][B][N]
    const float* W;                // [E][K][N]
    float wd;
    const float* zprev;            // DX: pre-activations of the previous layer [E][B][K]
    float* out;                    // DW: dW [E][K][N];  DX: delta of the previous layer [E][B][K]
    float* dbias;                  // DW: [E][N]
};
struct GemmBatch {
    int e0, ne, nl;                // blockIdx.z = layer slot * ne + member
    int64_t B;
    float* loss_terms;
    GemmLayer l[MB_MAX_LAYERS];
};

template <bool DX>
__global__ void __launch_bounds__(256)
k_train_gemm(GemmBatch gb) {
    __shared__ __align__(16) float As[kGR][kGLD];
    __shared__ __align__(16) float Bs[kGR][kGLD];
    __shared__ float red[8];
    const GemmLayer& ly = gb.l[blockIdx.z / gb.ne];
    const int e = gb.e0 + (int)(blockIdx.z % gb.ne);
    const int64_t B = gb.B;
    const int K = ly.K, N = ly.N;
    const float* __restrict__ hin = ly.hin;
    const int hin_is_z = ly.hin_is_z;
    const float* __restrict__ g = ly.g;
    const float* __restrict__ W = ly.W;
    const float wd = ly.wd;
    const float* __restrict__ zprev = ly.zprev;
    float* __restrict__ out = ly.out;
    float* __restrict__ dbias = ly.dbias;
    float* __restrict__ loss_terms = gb.loss_terms;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int M = DX ? (int)B : K;        // output rows
    const int NN = DX ? K : N;            // output cols
    const int R = DX ? N : (int)B;        // reduction length
    const int m0 = blockIdx.y * kGT, n0 = blockIdx.x * kGT;
    if (m0 >= M || n0 >= NN) return;      // batched launches size the grid for the widest layer
    const float* ge = g + (int64_t)e * B * N;
    const float* he = DX ? nullptr : hin + (int64_t)e * B * K;
    const float* We = W + (int64_t)e * K * N;
    float ra[kGPer], rb[kGPer];
    // DX: both operands have the REDUCTION index contiguous in memory (g[b][n], W[k][n]):
    //     thread -> output index tid/4, 8 consecutive reduction indices (stored transposed)
    // DW: both operands have the OUTPUT index contiguous (h[b][k], g[b][n]):
    //     thread -> reduction index tid/8, 8 consecutive output indices
    const int xo = tid >> 2, xr = (tid & 3) * kGPer;
    const int wr = tid >> 3, wo = (tid & 7) * kGPer;

    // 16-byte loads when K and N are multiples of 4 (every row then starts 16-byte aligned)
    const bool vec = (K & 3) == 0 && (N & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(ge) | reinterpret_cast<uintptr_t>(We) |
                       reinterpret_cast<uintptr_t>(he)) & 15) == 0;
    auto ld4 = [](const float* p, bool ok, float* dst) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(reinterpret_cast<const float4*>(p));
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    };
    auto fetch = [&](int r0) {
        if (DX) {
            if (vec) {
#pragma unroll
                for (int h = 0; h < kGPer / 4; ++h) {
                    const int r = r0 + xr + 4 * h;
                    ld4(ge + (int64_t)(m0 + xo) * N + r, r < R && m0 + xo < M, ra + 4 * h);
                    ld4(We + (int64_t)(n0 + xo) * N + r, r < R && n0 + xo < NN, rb + 4 * h);
                }
                return;
            }
#pragma unroll
            for (int c = 0; c < kGPer; ++c) {
                const int r = r0 + xr + c;
                const bool rin = r < R;
                ra[c] = (rin && m0 + xo < M) ? ge[(int64_t)(m0 + xo) * N + r] : 0.f;
                rb[c] = (rin && n0 + xo < NN) ? We[(int64_t)(n0 + xo) * N + r] : 0.f;
            }
        } else {
            const bool rin = r0 + wr < R;
            if (vec) {
#pragma unroll
                for (int h = 0; h < kGPer / 4; ++h) {
                    ld4(he + (int64_t)(r0 + wr) * K + m0 + wo + 4 * h, rin && m0 + wo + 4 * h < M, ra + 4 * h);
                    ld4(ge + (int64_t)(r0 + wr) * N + n0 + wo + 4 * h, rin && n0 + wo + 4 * h < NN, rb + 4 * h);
                }
                return;
            }
#pragma unroll
            for (int c = 0; c < kGPer; ++c) {
                ra[c] = (rin && m0 + wo + c < M) ? he[(int64_t)(r0 + wr) * K + m0 + wo + c] : 0.f;
                rb[c] = (rin && n0 + wo + c < NN) ? ge[(int64_t)(r0 + wr) * N + n0 + wo + c] : 0.f;
            }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;

    fetch(0);
    for (int r0 = 0; r0 < R; r0 += kGR) {
#pragma unroll
        for (int c = 0; c < kGPer; ++c) {
            if (DX) {
                As[xr + c][xo] = ra[c];
                Bs[xr + c][xo] = rb[c];
            } else {
                As[wr][wo + c] = hin_is_z ? swishf_(ra[c]) : ra[c];
                Bs[wr][wo + c] = rb[c];
            }
        }
        __syncthreads();
        if (r0 + kGR < R) fetch(r0 + kGR);
#pragma unroll
        for (int r = 0; r < kGR; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (!DX && blockIdx.y == 0 && tid < kGT) {
#pragma unroll
            for (int r = 0; r < kGR; ++r) bsum += Bs[r][tid];
        }
        __syncthreads();
    }
    float l2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mi = m0 + ty * 4 + i;
        if (mi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nj = n0 + tx * 4 + j;
            if (nj >= NN) continue;
            if (DX) {
                const float z = zprev[((int64_t)e * B + mi) * K + nj];
                const float sg = sigmoidf_(z);
                out[((int64_t)e * B + mi) * K + nj] = acc[i][j] * (sg * (1.f + z * (1.f - sg)));
            } else {
                const float w = We[(int64_t)mi * N + nj];
                out[(int64_t)e * K * N + (int64_t)mi * N + nj] = acc[i][j] + wd * w;
                l2 += 0.5f * wd * w * w;
            }
        }
    }
    if (!DX) {
        if (blockIdx.y == 0 && tid < kGT && n0 + tid < N) dbias[(int64_t)e * N + n0 + tid] = bsum;
        for (int o = 16; o; o >>= 1) l2 += __shfl_xor_sync(0xffffffffu, l2, o);
        if ((tid & 31) == 0) red[tid >> 5] = l2;
        __syncthreads();
        if (tid == 0) {
            float sacc = 0.f;
            for (int w = 0; w < 8; ++w) sacc += red[w];
            atomicAdd(loss_terms + 1, sacc);
        }
    }
}

// ---- Adam (torch.optim.Adam defaults, no weight decay / amsgrad) ----------------------------
struct AdamSegs { int n; int64_t off[2 * MB_MAX_LAYERS + 2]; int64_t len[2 * MB_MAX_LAYERS + 2]; };

// Advances the device-resident step count and derives Adam's two bias-correction scalars from it
// ONCE (double-precision pow is slow on this part; every k_adam block used to redo it).
__global__ void k_step_inc(int32_t* step_counter, float* scal, float lr, float beta1, float beta2) {
    pdl_trigger();
    pdl_wait();
    *step_counter += 1;
    const double t = (double)*step_counter;
    scal[0] = (float)((double)lr / (1.0 - pow((double)beta1, t)));
    scal[1] = (float)sqrt(1.0 - pow((double)beta2, t));
}

// scal != nullptr: bias corrections come from the device-resident step counter (graph replay)
__global__ void k_adam(AdamSegs segs, float* __restrict__ p, float* __restrict__ m1,
                       float* __restrict__ m2, const float* __restrict__ g, float beta1,
                       float beta2, float eps, float step_size, float bc2_sqrt,
                       const float* __restrict__ scal) {
    pdl_trigger();
    pdl_wait();
    if (scal != nullptr) {                   // graph replay: scalars written by k_step_inc
        step_size = scal[0];
        bc2_sqrt = scal[1];
    }
    const int s = blockIdx.y;
    const int64_t off = segs.off[s], len = segs.len[s];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = off + i;
        const float gr = g[k];
        const float a = m1[k] * beta1 + gr * (1.f - beta1);
        const float v = m2[k] * beta2 + (gr * gr) * (1.f - beta2);
        m1[k] = a;
        m2[k] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p[k] = p[k] - step_size * (a / denom);
    }
}

// activation slot == parameter member for training (members e0 .. e0+ne-1)
static TgBatch tg_members(int e0, int ne) {
    TgBatch tb{};
    tb.nmem = ne;
    tb.slot0 = e0;
    for (int i = 0; i < ne; ++i) tb.member[i] = e0 + i;
    for (int i = 0; i < kTgMaxOps; ++i) { tb.op[i].a_act = -1; tb.op[i].a_ones_row = -1; tb.op[i].out_act = -1; }
    return tb;
}

// MB_TRAIN_GEMM=simt selects the fp32 SIMT GEMMs (layer_gemm.cuh / k_train_gemm) for A/B timing
static bool train_gemm_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MB_TRAIN_GEMM");
        v = (e != nullptr && e[0] == 's') ? 0 : 1;
    }
    return v == 1;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return MB_ECUDA; }
    return MB_OK;
}

int launch_pe_train(const PeDev& m, void* tc_pack, const mb_train_cfg& cfg, float* params, float* exp_avg,
                    float* exp_avg_sq, int64_t step, int32_t* step_counter, const float* obs, const float* act,
                    const float* delta, const float* reward, const float* done, int64_t B,
                    int e0, int e1, float* grads_out, float* ensemble_loss, float* loss_terms,
                    void* ws, cudaStream_t s) {
    using Tile = SimtTile64;
    const mb_mlp_desc& d = m.mlp;
    const int L = d.num_layers, E = d.ensemble_size, D = d.sizes[L] / 2, ne = e1 - e0;
    TrainWs w = carve(d, B, ws);
    float* grads = grads_out ? grads_out : w.grads;

    float* gcur = w.gA;
    float* gnext = w.gB;
    const bool use_tc = train_gemm_tc() && kept_deltas(d, B) && d.activation == MB_ACT_SWISH;
    if (B * ne <= 65536) {
        // small batch: one launch per layer spreads the work over all SMs
        {
            const int64_t total = (int64_t)ne * B * d.sizes[0];
            launch_pdl(k_train_prep, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, m, e0, ne, obs, act, B,
                       w.x0, grads, ensemble_loss, loss_terms);
            MB_CUDA_LAUNCH_CHECK("k_train_prep");
        }
        // One launch for the whole forward when the fused tile kernel covers the shape: the
        // rollout kernel's pipeline in forward mode, saving pre-activations and the raw head row
        const bool fused_fwd = use_tc && tc_pack != nullptr && tc_rollout_supported(d);
        if (fused_fwd) {
            int rc = launch_pe_train_fwd_tc(m, tc_pack, obs, act, B, e0, ne, w.z, gnext, s);
            if (rc) return rc;
        }
        for (int l = 0; l < L && !fused_fwd; ++l) {
            if (use_tc) {
                TgBatch tb = tg_members(e0, ne);
                tb.nops = 1;
                TgOp& o = tb.op[0];
                o.mode = TG_FWD;
                o.M = (int)B; o.N = d.sizes[l + 1]; o.R = d.sizes[l];
                o.a = l == 0 ? w.x0 : w.z[l - 1]; o.lda = o.R; o.a_ss = (int64_t)B * o.R;
                o.a_act = l == 0 ? -1 : d.activation;         // layers consume stored PRE-activations
                o.b = m.params + w_offset(d, l); o.ldb = o.N; o.b_ms = (int64_t)o.R * o.N;
                o.bias = m.params + b_offset(d, l); o.bias_ms = o.N;
                o.out = l + 1 < L ? w.z[l] : gnext; o.ldo = o.N; o.out_ss = (int64_t)B * o.N;
                int rc = launch_tc_gemm(tb, s);
                if (rc) return rc;
                continue;
            }
            LayerArgs a;
            a.B = (int)B; a.K = d.sizes[l]; a.N = d.sizes[l + 1];
            a.A = l == 0 ? w.x0 : w.z[l - 1];
            a.a_slot_stride = (int64_t)B * a.K;
            a.W = m.params + w_offset(d, l);
            a.bias = m.params + b_offset(d, l);
            a.out = l + 1 < L ? w.z[l] : gnext;              // raw head output parks in the spare g buffer
            a.a_act = l == 0 ? -1 : d.activation;            // layers consume stored PRE-activations
            a.out_act = -1;
            a.slot0 = e0;
            a.nmem = ne;
            for (int i = 0; i < ne; ++i) a.member[i] = e0 + i;
            int rc = launch_layer(a, nullptr, s);
            if (rc) return rc;
        }
        dim3 gl((unsigned)((B + kLossRows - 1) / kLossRows), (unsigned)ne);
        launch_pdl(k_train_loss, gl, dim3(256), 0, s, m, cfg, e0, (const float*)gnext, delta, reward, done, B, gcur,
                   grads, ensemble_loss, loss_terms);
        MB_CUDA_LAUNCH_CHECK("k_train_loss");
    } else {
        // large batch: fused tile kernel keeps the activations on chip
        cudaMemsetAsync(ensemble_loss + e0, 0, sizeof(float) * ne, s);
        cudaMemsetAsync(loss_terms, 0, sizeof(float) * 3, s);
        cudaMemsetAsync(grads + logstd_offset(d, 0) + (int64_t)e0 * D, 0, sizeof(float) * ne * D, s);
        cudaMemsetAsync(grads + logstd_offset(d, 1) + (int64_t)e0 * D, 0, sizeof(float) * ne * D, s);
        int rc = set_smem(k_train_fwd, Tile::kSmemBytes);
        if (rc) return rc;
        ZPtrs zs;
        for (int l = 0; l < MB_MAX_LAYERS; ++l) zs.p[l] = w.z[l];
        dim3 gf((unsigned)((B + Tile::TM - 1) / Tile::TM), (unsigned)ne);
        k_train_fwd<<<gf, kSimtThreads, Tile::kSmemBytes, s>>>(m, cfg, e0, obs, act, delta, reward, done, B,
                                                               w.x0, zs, gcur, grads, ensemble_loss,
                                                               loss_terms);
        MB_CUDA_LAUNCH_CHECK("k_train_fwd");
    }
    if (use_tc) {
        // tcgen05 path: dX chain layer by layer, then every weight gradient in one launch
        TgBatch dwb = tg_members(e0, ne);
        for (int l = L - 1; l >= 0; --l) {
            const int K = d.sizes[l], N = d.sizes[l + 1];
            TgOp& o = dwb.op[dwb.nops++];
            o.mode = TG_DW;
            o.M = K + 1; o.N = N; o.R = (int)B;
            o.a = l == 0 ? w.x0 : w.z[l - 1]; o.lda = K; o.a_ss = (int64_t)B * K; o.a_trans = 1;
            o.a_act = l == 0 ? -1 : d.activation;
            o.a_ones_row = K;                                   // all-ones row: db = column sums of delta
            o.b = gcur; o.ldb = N; o.b_ss = (int64_t)B * N;
            o.w = m.params + w_offset(d, l); o.w_ms = (int64_t)K * N; o.wd = cfg.weight_decay[l];
            o.out = grads + w_offset(d, l); o.ldo = N; o.out_ms = (int64_t)K * N;
            o.out_ones = grads + b_offset(d, l); o.out_ones_ms = N;
            o.l2_acc = loss_terms + 1;
            if (l > 0) {
                TgBatch xb = tg_members(e0, ne);
                xb.nops = 1;
                TgOp& x = xb.op[0];
                x.mode = TG_DX;
                x.M = (int)B; x.N = K; x.R = N;
                x.a = gcur; x.lda = N; x.a_ss = (int64_t)B * N;
                x.b = m.params + w_offset(d, l); x.ldb = N; x.b_ms = (int64_t)K * N; x.b_trans = 1;
                x.z = w.z[l - 1]; x.ldz = K; x.z_ss = (int64_t)B * K;
                x.out = w.gl[l - 1]; x.ldo = K; x.out_ss = (int64_t)B * K;
                int rc = launch_tc_gemm(xb, s);
                if (rc) return rc;
                gcur = w.gl[l - 1];
            }
        }
        int rc = launch_tc_gemm(dwb, s);
        if (rc) return rc;
    } else {
        const bool batched_dw = kept_deltas(d, B) && B * ne <= kSmallBatchRows;
        GemmBatch dw{};
        dw.e0 = e0; dw.ne = ne; dw.nl = 0; dw.B = B; dw.loss_terms = loss_terms;
        int maxK = 0, maxN = 0;
        for (int l = L - 1; l >= 0; --l) {
            const int K = d.sizes[l], N = d.sizes[l + 1];
            const float* Wl = m.params + w_offset(d, l);
            GemmLayer gl{};
            gl.K = K; gl.N = N; gl.hin = l == 0 ? w.x0 : w.z[l - 1]; gl.hin_is_z = l > 0; gl.g = gcur; gl.W = Wl;
            gl.wd = cfg.weight_decay[l]; gl.out = grads + w_offset(d, l); gl.dbias = grads + b_offset(d, l);
            if (batched_dw) {
                dw.l[dw.nl++] = gl;
                maxK = K > maxK ? K : maxK;
                maxN = N > maxN ? N : maxN;
            } else {
                GemmBatch one{};
                one.e0 = e0; one.ne = ne; one.nl = 1; one.B = B; one.loss_terms = loss_terms; one.l[0] = gl;
                dim3 gw((unsigned)((N + kGT - 1) / kGT), (unsigned)((K + kGT - 1) / kGT), (unsigned)ne);
                k_train_gemm<false><<<gw, 256, 0, s>>>(one);
                MB_CUDA_LAUNCH_CHECK("k_train_gemm<DW>");
            }
            if (l > 0) {
                // with kept deltas every layer's delta has its own buffer (the weight-gradient GEMMs
                // read them later); otherwise two buffers ping-pong
                float* dst = batched_dw ? w.gl[l - 1] : gnext;
                GemmBatch dx{};
                dx.e0 = e0; dx.ne = ne; dx.nl = 1; dx.B = B; dx.loss_terms = nullptr;
                dx.l[0].K = K; dx.l[0].N = N; dx.l[0].g = gcur; dx.l[0].W = Wl; dx.l[0].zprev = w.z[l - 1];
                dx.l[0].out = dst;
                dim3 gx((unsigned)((K + kGT - 1) / kGT), (unsigned)((B + kGT - 1) / kGT), (unsigned)ne);
                k_train_gemm<true><<<gx, 256, 0, s>>>(dx);
                MB_CUDA_LAUNCH_CHECK("k_train_gemm<DX>");
                if (batched_dw) gcur = dst;
                else { float* t = gcur; gcur = gnext; gnext = t; }
            }
        }
        if (batched_dw) {
            // all weight gradients (layers x members independent GEMMs) in one launch
            dim3 gw((unsigned)((maxN + kGT - 1) / kGT), (unsigned)((maxK + kGT - 1) / kGT), (unsigned)(dw.nl * ne));
            k_train_gemm<false><<<gw, 256, 0, s>>>(dw);
            MB_CUDA_LAUNCH_CHECK("k_train_gemm<DW all layers>");
        }
    }
    if (params != nullptr) {
        AdamSegs segs;
        segs.n = 0;
        int64_t maxlen = 0;
        for (int l = 0; l < L; ++l) {
            const int64_t wn = (int64_t)d.sizes[l] * d.sizes[l + 1], bn = d.sizes[l + 1];
            segs.off[segs.n] = w_offset(d, l) + e0 * wn; segs.len[segs.n++] = ne * wn;
            segs.off[segs.n] = b_offset(d, l) + e0 * bn; segs.len[segs.n++] = ne * bn;
            maxlen = wn * ne > maxlen ? wn * ne : maxlen;
        }
        segs.off[segs.n] = logstd_offset(d, 0) + (int64_t)e0 * D; segs.len[segs.n++] = (int64_t)ne * D;
        segs.off[segs.n] = logstd_offset(d, 1) + (int64_t)e0 * D; segs.len[segs.n++] = (int64_t)ne * D;
        const double hstep = step >= 1 ? (double)step : 1.0;
        const double bc1 = 1.0 - pow((double)cfg.beta1, hstep);
        const double bc2 = 1.0 - pow((double)cfg.beta2, hstep);
        const int64_t nb = (maxlen + 255) / 256;
        dim3 ga((unsigned)(nb < 1184 ? nb : 1184), (unsigned)segs.n);
        float* scal = nullptr;
        if (step_counter) {
            scal = w.grads + round_up4(param_count(d, true));          // inside the workspace's tail slack
            launch_pdl(k_step_inc, dim3(1), dim3(1), 0, s, step_counter, scal, cfg.lr, cfg.beta1, cfg.beta2);
            MB_CUDA_LAUNCH_CHECK("k_step_inc");
        }
        launch_pdl(k_adam, ga, dim3(256), 0, s, segs, params, exp_avg, exp_avg_sq, (const float*)grads, cfg.beta1,
                   cfg.beta2, cfg.eps, (float)(cfg.lr / bc1), (float)sqrt(bc2), (const float*)scal);
        MB_CUDA_LAUNCH_CHECK("k_adam");
    }
    (void)E;
    return MB_OK;
}

// ---- critic update as a chain of launches (widths above 256: TQC's 3 x 512, configs/tqc.json) --
// Same maths as the persistent kernel's critic mode (train_fused.cu) for networks it does not cover, plus the
// quantile-Huber head of TQC (tqc_agent.py:5-14).  Input [obs | act] is shared by the members
// (ensemble_value.py:59-63, distributional_value.py:95-101); every member sees the same target.
struct CriticWs {
    float* x0;                     // [B][IN]
    float* z[MB_MAX_LAYERS];       // [E][B][sizes[l+1]] pre-activations, l < L-1
    float* out;                    // [E][B][Nout]
    float* g[MB_MAX_LAYERS];       // dL/d(output of layer l) [E][B][sizes[l+1]]
    float* grads;
};

size_t critic_chain_workspace_bytes(const mb_mlp_desc& d, int64_t B) {
    const int64_t E = d.ensemble_size;
    int64_t f = round_up4(B * d.sizes[0]);
    for (int l = 0; l < d.num_layers; ++l) f += 2 * round_up4(E * B * d.sizes[l + 1]);
    f += round_up4(param_count(d, false));
    return (size_t)f * sizeof(float) + 256;
}

static CriticWs critic_carve(const mb_mlp_desc& d, int64_t B, void* ws) {
    CriticWs w{};
    const int64_t E = d.ensemble_size;
    const int L = d.num_layers;
    float* p = reinterpret_cast<float*>(ws);
    w.x0 = p; p += round_up4(B * d.sizes[0]);
    for (int l = 0; l < L; ++l) {
        float* t = p; p += round_up4(E * B * d.sizes[l + 1]);
        if (l + 1 < L) w.z[l] = t; else w.out = t;
    }
    for (int l = 0; l < L; ++l) { w.g[l] = p; p += round_up4(E * B * d.sizes[l + 1]); }
    w.grads = p;
    return w;
}

__global__ void k_critic_prep(const float* __restrict__ obs, const float* __restrict__ act, int O, int A, int64_t B,
                              float* __restrict__ x0, float* __restrict__ ensemble_loss, int E,
                              float* __restrict__ loss_terms) {
    pdl_trigger();
    pdl_wait();
    const int IN = O + A;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < E) ensemble_loss[t] = 0.f;
    if (t < 3) loss_terms[t] = 0.f;
    if (t >= B * IN) return;
    const int64_t b = t / IN;
    const int j = (int)(t - b * IN);
    x0[t] = j < O ? obs[b * O + j] : act[b * A + (j - O)];
}

// One thread per (member, row, output): loss share and dL/d(output).
//   CRITIC_HEAD_MSE:      F.mse_loss against the shared target [B][Q] (ddpg_agent.py:138-149)
//   CRITIC_HEAD_QUANTILE: quantile_huber_loss_f against target atoms [B][T] (tqc_agent.py:5-14)
__global__ void k_critic_head(int kind, const float* __restrict__ out, const float* __restrict__ target, int T,
                              int64_t B, int E, int Q, float clip_error, float* __restrict__ g,
                              float* __restrict__ ensemble_loss, float* __restrict__ loss_terms) {
    pdl_trigger();
    pdl_wait();
    const int64_t per = B * Q, total = per * E;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float part = 0.f;
    int e = -1;
    if (t < total) {
        e = (int)(t / per);
        const int64_t r = t - (int64_t)e * per;
        const int64_t b = r / Q;
        const int j = (int)(r - b * Q);
        const float q = out[t];
        if (kind == CRITIC_HEAD_MSE) {
            const float wgt = 1.f / ((float)E * (float)B * (float)Q);
            float diff = q - target[b * Q + j];
            if (clip_error > 0.f) diff = fminf(fmaxf(diff, -clip_error), clip_error);
            part = diff * diff * wgt;
            g[t] = 2.f * diff * wgt;
        } else {
            const float wgt = 1.f / ((float)E * (float)B * (float)Q * (float)T);
            const float tau = ((float)j + 0.5f) / (float)Q;
            const float* tg = target + b * T;
            float gr = 0.f;
            for (int k = 0; k < T; ++k) {
                const float dl = __ldg(tg + k) - q;
                const float ad = fabsf(dl);
                const float hub = ad > 1.f ? ad - 0.5f : 0.5f * dl * dl;
                const float w = fabsf(tau - (dl < 0.f ? 1.f : 0.f));
                part += w * hub;
                gr -= w * fminf(fmaxf(dl, -1.f), 1.f);       // d huber / d q = -clip(delta, -1, 1)
            }
            part *= wgt;
            g[t] = gr * wgt;
        }
    }
    // warps never straddle two members when B*Q is a multiple of 32; otherwise one atomic per thread
    const bool uniform = (per & 31) == 0;
    if (uniform) {
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0 && e >= 0) { atomicAdd(ensemble_loss + e, part); atomicAdd(loss_terms, part); }
    } else if (e >= 0) {
        atomicAdd(ensemble_loss + e, part);
        atomicAdd(loss_terms, part);
    }
}

// ptu.soft_update_from_to (torch_modules/utils.py:151-155) after the optimizer step
__global__ void k_soft_update(const float* __restrict__ p, float* __restrict__ tp, float tau, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        tp[i] = tp[i] * (1.f - tau) + p[i] * tau;
}

int launch_critic_train_chain(const mb_mlp_desc& d, int head_kind, float* params, float* exp_avg, float* exp_avg_sq,
                              int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                              const float* obs, const float* act, int O, int A, const float* target, int T,
                              int64_t B, float clip_error, float* target_params, float tau, float* ensemble_loss,
                              float* loss_terms, void* ws, cudaStream_t s) {
    const int L = d.num_layers, E = d.ensemble_size, Q = d.sizes[L];
    if (E > 32) { set_error("critic chain: ensemble_size %d > 32", E); return MB_EUNSUPPORTED; }
    CriticWs w = critic_carve(d, B, ws);
    {
        int64_t total = B * d.sizes[0];
        total = total < E ? E : total;
        launch_pdl(k_critic_prep, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, obs, act, O, A, B, w.x0,
                   ensemble_loss, E, loss_terms);
        MB_CUDA_LAUNCH_CHECK("k_critic_prep");
    }
    for (int l = 0; l < L; ++l) {
        TgBatch tb = tg_members(0, E);
        tb.nops = 1;
        TgOp& o = tb.op[0];
        o.mode = TG_FWD;
        o.M = (int)B; o.N = d.sizes[l + 1]; o.R = d.sizes[l];
        o.a = l == 0 ? w.x0 : w.z[l - 1]; o.lda = o.R; o.a_ss = l == 0 ? 0 : (int64_t)B * o.R;
        o.a_act = l == 0 ? -1 : d.activation;
        o.b = params + w_offset(d, l); o.ldb = o.N; o.b_ms = (int64_t)o.R * o.N;
        o.bias = params + b_offset(d, l); o.bias_ms = o.N;
        o.out = l + 1 < L ? w.z[l] : w.out; o.ldo = o.N; o.out_ss = (int64_t)B * o.N;
        int rc = launch_tc_gemm(tb, s);
        if (rc) return rc;
    }
    {
        const int64_t total = (int64_t)E * B * Q;
        launch_pdl(k_critic_head, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, head_kind,
                   (const float*)w.out, target, T, B, E, Q, clip_error, w.g[L - 1], ensemble_loss, loss_terms);
        MB_CUDA_LAUNCH_CHECK("k_critic_head");
    }
    TgBatch dwb = tg_members(0, E);
    for (int l = L - 1; l >= 0; --l) {
        const int K = d.sizes[l], N = d.sizes[l + 1];
        TgOp& o = dwb.op[dwb.nops++];
        o.mode = TG_DW;
        o.M = K + 1; o.N = N; o.R = (int)B;
        o.a = l == 0 ? w.x0 : w.z[l - 1]; o.lda = K; o.a_ss = l == 0 ? 0 : (int64_t)B * K; o.a_trans = 1;
        o.a_act = l == 0 ? -1 : d.activation;
        o.a_ones_row = K;
        o.b = w.g[l]; o.ldb = N; o.b_ss = (int64_t)B * N;
        o.w = params + w_offset(d, l); o.w_ms = (int64_t)K * N; o.wd = 0.f;
        o.out = w.grads + w_offset(d, l); o.ldo = N; o.out_ms = (int64_t)K * N;
        o.out_ones = w.grads + b_offset(d, l); o.out_ones_ms = N;
        o.l2_acc = nullptr;
        if (l > 0) {
            TgBatch xb = tg_members(0, E);
            xb.nops = 1;
            TgOp& x = xb.op[0];
            x.mode = TG_DX;
            x.M = (int)B; x.N = K; x.R = N;
            x.a = w.g[l]; x.lda = N; x.a_ss = (int64_t)B * N;
            x.b = params + w_offset(d, l); x.ldb = N; x.b_ms = (int64_t)K * N; x.b_trans = 1;
            x.z = w.z[l - 1]; x.ldz = K; x.z_ss = (int64_t)B * K; x.z_act = d.activation;
            x.out = w.g[l - 1]; x.ldo = K; x.out_ss = (int64_t)B * K;
            int rc = launch_tc_gemm(xb, s);
            if (rc) return rc;
        }
    }
    int rc = launch_tc_gemm(dwb, s);
    if (rc) return rc;
    AdamSegs segs;
    segs.n = 0;
    const int64_t n = param_count(d, false);
    int64_t maxlen = 0;
    for (int l = 0; l < L; ++l) {
        const int64_t wn = (int64_t)E * d.sizes[l] * d.sizes[l + 1], bn = (int64_t)E * d.sizes[l + 1];
        segs.off[segs.n] = w_offset(d, l); segs.len[segs.n++] = wn;
        segs.off[segs.n] = b_offset(d, l); segs.len[segs.n++] = bn;
        maxlen = wn > maxlen ? wn : maxlen;
    }
    const double hstep = (double)(adam_steps_done + 1);
    const double bc1 = 1.0 - pow((double)beta1, hstep), bc2 = 1.0 - pow((double)beta2, hstep);
    const int64_t nb = (maxlen + 255) / 256;
    launch_pdl(k_adam, dim3((unsigned)(nb < 1184 ? nb : 1184), (unsigned)segs.n), dim3(256), 0, s, segs, params, exp_avg, exp_avg_sq,
               (const float*)w.grads, beta1, beta2, eps, (float)(lr / bc1), (float)sqrt(bc2), (const float*)nullptr);
    MB_CUDA_LAUNCH_CHECK("k_adam");
    if (target_params != nullptr) {
        launch_pdl(k_soft_update, dim3((unsigned)(nb < 1184 ? nb : 1184)), dim3(256), 0, s, (const float*)params,
                   target_params, tau, n);
        MB_CUDA_LAUNCH_CHECK("k_soft_update");
    }
    return MB_OK;
}

// ---- actor + entropy-coefficient update as a chain of launches -------------------------
// SACAgent.train_from_torch_batch, "update actor and alpha" (sac_agent.py:170-195):
//   new_action, info = policy.action(obs, reparameterize=True, return_log_prob=True)   (gaussian.py:55-97)
//   alpha:  res = log_prob.mean() + target_entropy;  loss = -res * exp(log_alpha);  Adam    (sac_agent.py:88-100)
//   policy: loss = alpha * log_prob.mean() - qf.value(obs, new_action, **current_v_pi_kwargs).mean()   (:102-118)
// written out by hand: policy forward (tensor-core layer GEMMs, pre-activations kept), sample + log-prob head,
// alpha step, critic forward on [obs | action], d loss / d q through the member reduction, the critic's dX chain
// down to its INPUT (no critic weight gradients), d loss / d (mean, log-std) through tanh and the log-prob, the
// policy's dW / dX chain and Adam.  With pre = mean + std * eps fixed by the noise draw:
//   d log_prob / d pre = 2 tanh(pre) (squash correction; the Normal term's pre and mean paths cancel),
//   d log_prob / d std = -1 / std.
struct ActorWs {
    float* pz[MB_MAX_LAYERS];      // policy pre-activations [B][w], l < Lp-1
    float* pout;                   // [B][2A]
    float* action;                 // [B][A]
    float* pg[MB_MAX_LAYERS];      // dL/d(output of policy layer l) [B][w_{l+1}]
    float* pgrads;                 // policy gradient blob
    float* x0;                     // [B][O+A]
    float* cz[MB_MAX_LAYERS];      // critic pre-activations [E][B][w], l < Lc-1
    float* cout;                   // [E][B][Q]
    float* cg[MB_MAX_LAYERS];      // dL/d(output of critic layer l) [E][B][w_{l+1}]
    float* gx0;                    // [E][B][O+A]
    float* scal;                   // [0] sum log_prob  [1] sum q  [2] alpha of the policy loss  [3] alpha before its step  [4] alpha loss
};

static ActorWs actor_carve(const mb_mlp_desc& pd, const mb_mlp_desc& cd, int64_t B, void* ws, size_t* bytes) {
    ActorWs w{};
    const int Lp = pd.num_layers, Lc = cd.num_layers;
    const int64_t E = cd.ensemble_size;
    float* p = reinterpret_cast<float*>(ws);
    float* const p0 = p;
    auto take = [&](int64_t n) { float* t = p; p += round_up4(n); return t; };
    w.scal = take(16);
    for (int l = 0; l + 1 < Lp; ++l) w.pz[l] = take(B * pd.sizes[l + 1]);
    w.pout = take(B * pd.sizes[Lp]);
    w.action = take(B * (pd.sizes[Lp] / 2));
    for (int l = 0; l < Lp; ++l) w.pg[l] = take(B * pd.sizes[l + 1]);
    w.pgrads = take(param_count(pd, false));
    w.x0 = take(B * cd.sizes[0]);
    for (int l = 0; l + 1 < Lc; ++l) w.cz[l] = take(E * B * cd.sizes[l + 1]);
    w.cout = take(E * B * cd.sizes[Lc]);
    for (int l = 0; l < Lc; ++l) w.cg[l] = take(E * B * cd.sizes[l + 1]);
    w.gx0 = take(E * B * cd.sizes[0]);
    if (bytes) *bytes = (size_t)(p - p0) * sizeof(float) + 256;
    return w;
}

size_t actor_workspace_bytes(const mb_mlp_desc& pd, const mb_mlp_desc& cd, int64_t B) {
    size_t n = 0;
    actor_carve(pd, cd, B, nullptr, &n);
    return n;
}

__device__ __forceinline__ float bounded_logstd(float raw, int mode, float& dls_draw) {
    if (mode == MB_LOGSTD_TANH) {            // (tanh + 1) * (max - min) / 2 + min, gaussian.py:17-30
        const float th = tanhf(raw);
        dls_draw = 6.f * (1.f - th * th);
        return (th + 1.f) * 6.f - 10.f;
    }
    if (mode == MB_LOGSTD_CLAMP) {
        dls_draw = (raw >= -10.f && raw <= 2.f) ? 1.f : 0.f;
        return fminf(fmaxf(raw, -10.f), 2.f);
    }
    dls_draw = 1.f;
    return raw;
}

// sample + log-prob (k_policy_head's arithmetic); the log-probs are summed over the batch into scal[0]
__global__ void k_actor_head_fwd(const float* __restrict__ out, const float* __restrict__ eps, int64_t B, int A,
                                 mb_policy_cfg cfg, float* __restrict__ action, float* __restrict__ scal) {
    pdl_trigger();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float lp = 0.f;
    if (t < B * A) {
        const int64_t b = t / A;
        const int j = (int)(t - b * A);
        float dd;
        const float mean = out[b * 2 * A + j] * cfg.mean_scale;
        const float ls = bounded_logstd(out[b * 2 * A + A + j] + cfg.logstd_extra_bias, cfg.bound_mode, dd);
        const float sd = expf(ls) * cfg.std_scale + cfg.min_std;
        const float e = eps[t];
        const float pre = mean + sd * e;
        action[t] = cfg.squashed ? tanhf(pre) : pre;
        lp = -0.5f * e * e - logf(sd) - 0.91893853320467274178f;
        if (cfg.squashed) lp -= 2.f * 0.69314718055994530942f - softplusf_(2.f * pre) - softplusf_(-2.f * pre);
    }
    for (int o = 16; o; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(scal, lp);
}

// entropy coefficient: compute_alpha_loss + alpha_optimizer.step (torch.optim.Adam on the scalar log_alpha)
__global__ void k_alpha_update(float* __restrict__ scal, int64_t B, float target_entropy, float alpha_fixed,
                               float* log_alpha, float* m1, float* m2, int64_t steps_done, float lr, float beta1,
                               float beta2, float eps) {
    pdl_trigger();
    pdl_wait();
    if (log_alpha == nullptr) { scal[2] = alpha_fixed; scal[3] = alpha_fixed; scal[4] = 0.f; return; }
    const float la = *log_alpha, alpha = expf(la);
    const float res = scal[0] / (float)B + target_entropy;
    const float g = -res * alpha;
    scal[3] = alpha;
    scal[4] = g;                                     // alpha_loss = (-res) * alpha, numerically its own gradient
    const double t = (double)(steps_done + 1);
    const float a = *m1 * beta1 + g * (1.f - beta1), v = *m2 * beta2 + g * g * (1.f - beta2);
    *m1 = a;
    *m2 = v;
    const float step_size = (float)((double)lr / (1.0 - pow((double)beta1, t)));
    const float denom = sqrtf(v) / (float)sqrt(1.0 - pow((double)beta2, t)) + eps;
    const float nla = la - step_size * (a / denom);
    *log_alpha = nla;
    scal[2] = expf(nla);                              // _get_alpha() after the step (sac_agent.py:66-72,177-188)
}

struct ActorMembers { int n; int idx[32]; };

// q_pi of a row = reduction of the listed members' outputs (ensemble_value.py:91-96; TQC: the mean of all atoms,
// distributional_value.py:51-55), summed into scal[1]; seed of the backward pass: d(-mean_b q_pi) / d q[e][b][o]
__global__ void k_actor_seed(const float* __restrict__ cout, int E, int64_t B, int Q, int mode, ActorMembers mem,
                             float* __restrict__ g, float* __restrict__ scal) {
    pdl_trigger();
    pdl_wait();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float q = 0.f;
    if (b < B) {
        for (int e = 0; e < E; ++e)
            for (int o = 0; o < Q; ++o) g[((int64_t)e * B + b) * Q + o] = 0.f;
        const float invB = 1.f / (float)B;
        if (mode == MB_REDUCE_MEAN) {
            const float wgt = 1.f / (float)(mem.n * Q);
            for (int i = 0; i < mem.n; ++i)
                for (int o = 0; o < Q; ++o) {
                    const int64_t k = ((int64_t)mem.idx[i] * B + b) * Q + o;
                    q += cout[k] * wgt;
                    g[k] = -invB * wgt;
                }
        } else {
            int best = mem.idx[0];
            q = cout[((int64_t)best * B + b) * Q];
            for (int i = 1; i < mem.n; ++i) {
                const float v = cout[((int64_t)mem.idx[i] * B + b) * Q];
                if (mode == MB_REDUCE_MIN ? v < q : v > q) { q = v; best = mem.idx[i]; }
            }
            g[((int64_t)best * B + b) * Q] = -invB;
        }
    }
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(scal + 1, q);
}

// d loss / d (policy output): the members' input gradients are summed over e, the action columns go through the
// squash and the reparameterisation, the log-prob terms are added
__global__ void k_actor_head_bwd(const float* __restrict__ gx0, int E, int64_t B, int IN, int O, int A,
                                 const float* __restrict__ out, const float* __restrict__ eps, mb_policy_cfg cfg,
                                 const float* __restrict__ scal, float* __restrict__ gout) {
    pdl_trigger();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * A) return;
    const int64_t b = t / A;
    const int j = (int)(t - b * A);
    float ga = 0.f;
    for (int e = 0; e < E; ++e) ga += gx0[((int64_t)e * B + b) * IN + O + j];
    const float alpha_b = scal[2] / (float)B;
    float dls_draw;
    const float mean = out[b * 2 * A + j] * cfg.mean_scale;
    const float ls = bounded_logstd(out[b * 2 * A + A + j] + cfg.logstd_extra_bias, cfg.bound_mode, dls_draw);
    const float ex = expf(ls), sd = ex * cfg.std_scale + cfg.min_std;
    const float e = eps[t];
    const float pre = mean + sd * e;
    float dpre = ga;
    if (cfg.squashed) {
        const float a = tanhf(pre);
        dpre = ga * (1.f - a * a) + alpha_b * 2.f * a;
    }
    const float dsd = dpre * e - alpha_b / sd;
    gout[b * 2 * A + j] = dpre * cfg.mean_scale;
    gout[b * 2 * A + A + j] = dsd * cfg.std_scale * ex * dls_draw;
}

// info[0] policy loss, [1] q_pi mean, [2] entropy, [3] alpha before its step, [4] alpha loss, [5] alpha of the policy loss
__global__ void k_actor_info(const float* __restrict__ scal, int64_t B, float* __restrict__ info) {
    pdl_trigger();
    pdl_wait();
    const float lp = scal[0] / (float)B, q = scal[1] / (float)B;
    info[0] = scal[2] * lp - q;
    info[1] = q;
    info[2] = -lp;
    info[3] = scal[3];
    info[4] = scal[4];
    info[5] = scal[2];
}

int launch_actor_update(const mb_mlp_desc& pd, float* pparams, float* exp_avg, float* exp_avg_sq,
                        int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                        const mb_policy_cfg& cfg, const mb_mlp_desc& cd, const float* cparams, int reduce_mode,
                        const int32_t* members, int n_members, const float* obs, const float* noise, int O, int A,
                        int64_t B, float* log_alpha, float* la_m1, float* la_m2, float target_entropy,
                        float alpha_fixed, float* info, void* ws, cudaStream_t s) {
    const int Lp = pd.num_layers, Lc = cd.num_layers, E = cd.ensemble_size, Q = cd.sizes[Lc], IN = O + A;
    if (E > 32) { set_error("actor update: ensemble_size %d > 32", E); return MB_EUNSUPPORTED; }
    ActorWs w = actor_carve(pd, cd, B, ws, nullptr);
    ActorMembers mem{};
    mem.n = members ? n_members : E;
    for (int i = 0; i < mem.n; ++i) mem.idx[i] = members ? members[i] : i;
    cudaMemsetAsync(w.scal, 0, 16 * sizeof(float), s);
    const unsigned gBA = (unsigned)((B * A + 255) / 256);
    // ---- policy forward, pre-activations kept
    for (int l = 0; l < Lp; ++l) {
        TgBatch tb = tg_members(0, 1);
        tb.nops = 1;
        TgOp& o = tb.op[0];
        o.mode = TG_FWD;
        o.M = (int)B; o.N = pd.sizes[l + 1]; o.R = pd.sizes[l];
        o.a = l == 0 ? obs : w.pz[l - 1]; o.lda = o.R; o.a_ss = 0;
        o.a_act = l == 0 ? -1 : pd.activation;
        o.b = pparams + w_offset(pd, l); o.ldb = o.N; o.b_ms = (int64_t)o.R * o.N;
        o.bias = pparams + b_offset(pd, l); o.bias_ms = o.N;
        o.out = l + 1 < Lp ? w.pz[l] : w.pout; o.ldo = o.N; o.out_ss = (int64_t)B * o.N;
        int rc = launch_tc_gemm(tb, s);
        if (rc) return rc;
    }
    launch_pdl(k_actor_head_fwd, dim3(gBA), dim3(256), 0, s, (const float*)w.pout, noise, B, A, cfg, w.action, w.scal);
    MB_CUDA_LAUNCH_CHECK("k_actor_head_fwd");
    launch_pdl(k_alpha_update, dim3(1), dim3(1), 0, s, w.scal, B, target_entropy, alpha_fixed, log_alpha, la_m1, la_m2,
               adam_steps_done, lr, beta1, beta2, eps);
    MB_CUDA_LAUNCH_CHECK("k_alpha_update");
    // ---- critics on [obs | new action]
    {
        const int64_t total = B * IN;
        launch_pdl(k_critic_prep, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, obs, (const float*)w.action, O, A,
                   B, w.x0, w.scal + 8, 0, w.scal + 8);
        MB_CUDA_LAUNCH_CHECK("k_critic_prep");
    }
    for (int l = 0; l < Lc; ++l) {
        TgBatch tb = tg_members(0, E);
        tb.nops = 1;
        TgOp& o = tb.op[0];
        o.mode = TG_FWD;
        o.M = (int)B; o.N = cd.sizes[l + 1]; o.R = cd.sizes[l];
        o.a = l == 0 ? w.x0 : w.cz[l - 1]; o.lda = o.R; o.a_ss = l == 0 ? 0 : (int64_t)B * o.R;
        o.a_act = l == 0 ? -1 : cd.activation;
        o.b = cparams + w_offset(cd, l); o.ldb = o.N; o.b_ms = (int64_t)o.R * o.N;
        o.bias = cparams + b_offset(cd, l); o.bias_ms = o.N;
        o.out = l + 1 < Lc ? w.cz[l] : w.cout; o.ldo = o.N; o.out_ss = (int64_t)B * o.N;
        int rc = launch_tc_gemm(tb, s);
        if (rc) return rc;
    }
    launch_pdl(k_actor_seed, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, s, (const float*)w.cout, E, B, Q, reduce_mode,
               mem, w.cg[Lc - 1], w.scal);
    MB_CUDA_LAUNCH_CHECK("k_actor_seed");
    // ---- critic dX chain down to the input (no weight gradients: the critics are not updated here)
    for (int l = Lc - 1; l >= 0; --l) {
        const int K = cd.sizes[l], N = cd.sizes[l + 1];
        TgBatch xb = tg_members(0, E);
        xb.nops = 1;
        TgOp& x = xb.op[0];
        x.mode = TG_DX;
        x.M = (int)B; x.N = K; x.R = N;
        x.a = w.cg[l]; x.lda = N; x.a_ss = (int64_t)B * N;
        x.b = cparams + w_offset(cd, l); x.ldb = N; x.b_ms = (int64_t)K * N; x.b_trans = 1;
        if (l > 0) { x.z = w.cz[l - 1]; x.ldz = K; x.z_ss = (int64_t)B * K; x.z_act = cd.activation; }
        x.out = l > 0 ? w.cg[l - 1] : w.gx0; x.ldo = K; x.out_ss = (int64_t)B * K;
        int rc = launch_tc_gemm(xb, s);
        if (rc) return rc;
    }
    launch_pdl(k_actor_head_bwd, dim3(gBA), dim3(256), 0, s, (const float*)w.gx0, E, B, IN, O, A, (const float*)w.pout, noise,
               cfg, (const float*)w.scal, w.pg[Lp - 1]);
    MB_CUDA_LAUNCH_CHECK("k_actor_head_bwd");
    // ---- policy backward + Adam
    TgBatch dwb = tg_members(0, 1);
    for (int l = Lp - 1; l >= 0; --l) {
        const int K = pd.sizes[l], N = pd.sizes[l + 1];
        TgOp& o = dwb.op[dwb.nops++];
        o.mode = TG_DW;
        o.M = K + 1; o.N = N; o.R = (int)B;
        o.a = l == 0 ? obs : w.pz[l - 1]; o.lda = K; o.a_ss = 0; o.a_trans = 1;
        o.a_act = l == 0 ? -1 : pd.activation;
        o.a_ones_row = K;
        o.b = w.pg[l]; o.ldb = N; o.b_ss = (int64_t)B * N;
        o.w = pparams + w_offset(pd, l); o.w_ms = (int64_t)K * N; o.wd = 0.f;
        o.out = w.pgrads + w_offset(pd, l); o.ldo = N; o.out_ms = (int64_t)K * N;
        o.out_ones = w.pgrads + b_offset(pd, l); o.out_ones_ms = N;
        o.l2_acc = nullptr;
        if (l > 0) {
            TgBatch xb = tg_members(0, 1);
            xb.nops = 1;
            TgOp& x = xb.op[0];
            x.mode = TG_DX;
            x.M = (int)B; x.N = K; x.R = N;
            x.a = w.pg[l]; x.lda = N; x.a_ss = (int64_t)B * N;
            x.b = pparams + w_offset(pd, l); x.ldb = N; x.b_ms = (int64_t)K * N; x.b_trans = 1;
            x.z = w.pz[l - 1]; x.ldz = K; x.z_ss = (int64_t)B * K; x.z_act = pd.activation;
            x.out = w.pg[l - 1]; x.ldo = K; x.out_ss = (int64_t)B * K;
            int rc = launch_tc_gemm(xb, s);
            if (rc) return rc;
        }
    }
    int rc = launch_tc_gemm(dwb, s);
    if (rc) return rc;
    AdamSegs segs;
    segs.n = 0;
    int64_t maxlen = 0;
    for (int l = 0; l < Lp; ++l) {
        const int64_t wn = (int64_t)pd.sizes[l] * pd.sizes[l + 1], bn = pd.sizes[l + 1];
        segs.off[segs.n] = w_offset(pd, l); segs.len[segs.n++] = wn;
        segs.off[segs.n] = b_offset(pd, l); segs.len[segs.n++] = bn;
        maxlen = wn > maxlen ? wn : maxlen;
    }
    const double hstep = (double)(adam_steps_done + 1);
    const double bc1 = 1.0 - pow((double)beta1, hstep), bc2 = 1.0 - pow((double)beta2, hstep);
    const int64_t nb = (maxlen + 255) / 256;
    launch_pdl(k_adam, dim3((unsigned)(nb < 1184 ? nb : 1184), (unsigned)segs.n), dim3(256), 0, s, segs, pparams, exp_avg,
               exp_avg_sq, (const float*)w.pgrads, beta1, beta2, eps, (float)(lr / bc1), (float)sqrt(bc2),
               (const float*)nullptr);
    MB_CUDA_LAUNCH_CHECK("k_adam");
    if (info != nullptr) {
        launch_pdl(k_actor_info, dim3(1), dim3(1), 0, s, (const float*)w.scal, B, info);
        MB_CUDA_LAUNCH_CHECK("k_actor_info");
    }
    return MB_OK;
}

}  // namespace mb
