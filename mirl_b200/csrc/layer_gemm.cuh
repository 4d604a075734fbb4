// One ensemble layer as a batched SIMT fp32 GEMM:  out[slot][B][N] = f_out(f_in(A[slot]) * W[e] + b[e])
//
// Used where the batch is small (model training at 256 rows/member, REDQ / TQC targets at 256
// rows): there the fused tile kernels of simt_mlp.cuh leave most SMs idle (a 256-row batch is 4-8
// tiles), while a per-layer launch spreads one layer over (N/64) x (B/64) x members CTAs.
// Replaces MyEnsembleLinear.forward + activation (mirl/torch_modules/linear.py:143-149,
// torch_modules/utils.py:125-149); plain fp32 FFMA accumulation like the reference's SGEMM.
//
// Grids are at most ~1 CTA per SM here, so nothing hides a global-load round trip except the
// kernel itself: the reduction runs in chunks of 32 with the NEXT chunk's operands already in
// flight in registers while the current chunk is multiplied out of shared memory.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kLgT = 64;          // output tile edge
constexpr int kLgK = 32;          // reduction chunk
constexpr int kLgLD = kLgT + 4;
constexpr int kLgPer = kLgT * kLgK / 256;   // operand elements per thread per chunk (8)

struct LayerArgs {
    const float* A;          // [slots][B][K]; a_slot_stride == 0 -> one input shared by all members
    int64_t a_slot_stride;
    const float* W;          // blob + w_offset(layer): [E][K][N]
    const float* bias;       // blob + b_offset(layer): [E][N]
    float* out;              // [slots][B][N]
    int B, K, N;
    int a_act;               // applied to A on load (-1 none): lets a layer consume stored pre-activations
    int out_act;             // applied to the result (-1 none)
    int slot0;               // slot of member[0] (training: first member of the slice; critics: 0)
    int nmem;
    int member[32];          // member id per blockIdx.z
};

// Optional layer-0 input: concatenation [obs | act] with optional Normalizer.process
// (mirl/processors/normalizer.py:17-18); rows are per slot when per_slot != 0.
struct ConcatIn {
    const float* obs;
    const float* act;
    const float* norm;       // [obs_mean O | obs_std O | act_mean A | act_std A | ...] or nullptr
    int O, A, per_slot;
};

__device__ __forceinline__ float lg_act(float x, int act) { return act < 0 ? x : apply_act(x, act); }

template <bool CONCAT>
__global__ void __launch_bounds__(256)
k_layer_gemm(LayerArgs a, ConcatIn ci) {
    __shared__ __align__(16) float As[kLgK][kLgLD];
    __shared__ __align__(16) float Bs[kLgK][kLgLD];
    pdl_trigger();
    pdl_wait();
    const int z = blockIdx.z, e = a.member[z], slot = a.slot0 + z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kLgT, n0 = blockIdx.x * kLgT;
    const float* Ae = CONCAT ? nullptr : a.A + (int64_t)slot * a.a_slot_stride;
    const float* We = a.W + (int64_t)e * a.K * a.N;
    const int64_t crow0 = CONCAT && ci.per_slot ? (int64_t)slot * a.B : 0;
    // operand ownership: A -> row m = tid/4, 8 consecutive k;  W -> row r = tid/8, 8 consecutive n
    const int am = tid >> 2, ak = (tid & 3) * kLgPer;
    const int wr = tid >> 3, wn = (tid & 7) * kLgPer;
    float ra[kLgPer], rw[kLgPer];

    // 16-byte loads whenever both operands allow it (K, N multiples of 4 and 16-byte aligned
    // bases): two LDG.128 per operand per chunk instead of eight scalar loads
    const bool vec = !CONCAT && (a.K & 3) == 0 && (a.N & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(Ae) | reinterpret_cast<uintptr_t>(We)) & 15) == 0;
    auto fetch = [&](int r0) {
        const int m = m0 + am;
        if (vec) {
#pragma unroll
            for (int h = 0; h < kLgPer / 4; ++h) {
                const int r = r0 + ak + 4 * h;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < a.B && r < a.K) v = __ldg(reinterpret_cast<const float4*>(Ae + (int64_t)m * a.K + r));
                ra[4 * h + 0] = v.x; ra[4 * h + 1] = v.y; ra[4 * h + 2] = v.z; ra[4 * h + 3] = v.w;
                const int n = n0 + wn + 4 * h;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + wr < a.K && n < a.N) w = __ldg(reinterpret_cast<const float4*>(We + (int64_t)(r0 + wr) * a.N + n));
                rw[4 * h + 0] = w.x; rw[4 * h + 1] = w.y; rw[4 * h + 2] = w.z; rw[4 * h + 3] = w.w;
            }
            return;
        }
#pragma unroll
        for (int c = 0; c < kLgPer; ++c) {
            const int r = r0 + ak + c;
            float v = 0.f;
            if (m < a.B && r < a.K) {
                if (CONCAT) {
                    const int64_t row = crow0 + m;
                    if (r < ci.O) {
                        v = ci.obs[row * ci.O + r];
                        if (ci.norm) v = (v - ci.norm[r]) / (ci.norm[ci.O + r] + kNormEps);
                    } else {
                        const int k = r - ci.O;
                        v = ci.act[row * ci.A + k];
                        if (ci.norm) v = (v - ci.norm[2 * ci.O + k]) / (ci.norm[2 * ci.O + ci.A + k] + kNormEps);
                    }
                } else {
                    v = Ae[(int64_t)m * a.K + r];
                }
            }
            ra[c] = v;
        }
        const bool rin = r0 + wr < a.K;
#pragma unroll
        for (int c = 0; c < kLgPer; ++c)
            rw[c] = (rin && n0 + wn + c < a.N) ? We[(int64_t)(r0 + wr) * a.N + n0 + wn + c] : 0.f;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    fetch(0);
    for (int r0 = 0; r0 < a.K; r0 += kLgK) {
#pragma unroll
        for (int c = 0; c < kLgPer; ++c) {
            As[ak + c][am] = CONCAT ? ra[c] : lg_act(ra[c], a.a_act);       // transposed: As[k][m]
            Bs[wr][wn + c] = rw[c];
        }
        __syncthreads();
        if (r0 + kLgK < a.K) fetch(r0 + kLgK);             // in flight during the multiply below
#pragma unroll
        for (int r = 0; r < kLgK; ++r) {
            const float4 av4 = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 bv4 = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w}, bv[4] = {bv4.x, bv4.y, bv4.z, bv4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float* be = a.bias + (int64_t)e * a.N;
    float* oe = a.out + (int64_t)slot * a.B * a.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) oe[(int64_t)m * a.N + n] = lg_act(acc[i][j] + be[n], a.out_act);
        }
    }
}

inline int launch_layer(const LayerArgs& a, const ConcatIn* ci, cudaStream_t s) {
    dim3 grid((unsigned)((a.N + kLgT - 1) / kLgT), (unsigned)((a.B + kLgT - 1) / kLgT), (unsigned)a.nmem);
    if (ci) launch_pdl(k_layer_gemm<true>, grid, dim3(256), 0, s, a, *ci);
    else {
        ConcatIn none{};
        launch_pdl(k_layer_gemm<false>, grid, dim3(256), 0, s, a, none);
    }
    MB_CUDA_LAUNCH_CHECK("k_layer_gemm");
    return MB_OK;
}

}  // namespace mb
