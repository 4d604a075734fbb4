// SIMT fp32 ensemble-MLP tile engine.
//
// One CTA (256 threads) pushes a tile of TM rows through every layer of ONE ensemble member.
// Activations stay in shared memory ([TM][AS], row-major) for the whole stack; the member's
// weights stream from global/L2 in 16-row chunks through a cp.async double buffer.  Each thread
// owns an RM x (4*NQ) register tile (rows rg*RM.., column quads cg, cg+16, ...), so the inner
// loop is 1 LDS.128 of activations + NQ LDS.128 of weights per 4*RM*NQ FFMA.
//
// Replaces, per layer, the reference's `torch.stack(weights)` + `x.matmul(w) + b` + activation
// (mirl/torch_modules/linear.py:143-149, mlp.py:14-38, torch_modules/utils.py:125-130).
// Accumulation is plain fp32 FFMA in k order -- the same arithmetic class as the reference's
// fp32 SGEMM (no TF32).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kSimtThreads = 256;
constexpr int kKC = 16;  // weight rows per streamed chunk

template <int RM, int NQ>
struct SimtTile {
    static constexpr int TM = 16 * RM;     // rows per tile
    static constexpr int MAXW = 64 * NQ;   // widest layer supported
    static constexpr int AS = MAXW + 4;    // activation row stride (floats)
    static constexpr int kActFloats = TM * AS;
    static constexpr int kWFloats = kKC * MAXW;
    static constexpr size_t kSmemBytes = (size_t)(2 * kActFloats + 2 * kWFloats) * sizeof(float);

    // Stage rows [k0, k0+kKC) of W[K][N] into wdst[kKC][WS]; rows >= K and pad columns are zero.
    __device__ static void load_chunk(const float* __restrict__ Wg, int K, int N, int WS, int k0,
                                      float* wdst) {
        const int tid = threadIdx.x;
        const int rows = min(kKC, K - k0);
        if ((N & 3) == 0) {
            const int nvec = rows * N / 4;
            const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)k0 * N);
            float4* dst = reinterpret_cast<float4*>(wdst);
            for (int i = tid; i < nvec; i += kSimtThreads) cp_async16(dst + i, src + i);
            const int total = kKC * N / 4;
            for (int i = nvec + tid; i < total; i += kSimtThreads)
                dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = tid; i < kKC * WS; i += kSimtThreads) {
                int kk = i / WS, n = i - kk * WS;
                if (kk < rows && n < N) cp_async4(wdst + i, Wg + (size_t)(k0 + kk) * N + n);
                else wdst[i] = 0.f;
            }
        }
    }

    // out[TM][AS] = act(in[TM][.K] * W[K][N] + b).  zsave (optional) receives the
    // pre-activation rows: zsave[(row0 + r) * zstride + c], rows < nrows only.
    __device__ static void layer(const float* __restrict__ Wg, const float* __restrict__ bg, int K,
                                 int N, int act, const float* in, float* out, float* wbuf,
                                 float* zsave, int zstride, int nrows) {
        const int tid = threadIdx.x;
        const int cg = tid & 15, rg = tid >> 4;
        const int WS = (N + 3) & ~3;
        const int nquads = WS >> 2;
        const int K4 = (K + 3) & ~3;
        const int nchunks = (K + kKC - 1) / kKC;

        float acc[RM][NQ][4];
#pragma unroll
        for (int r = 0; r < RM; ++r)
#pragma unroll
            for (int j = 0; j < NQ; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][j][c] = 0.f;

        load_chunk(Wg, K, N, WS, 0, wbuf);
        cp_async_commit();
        for (int ch = 0; ch < nchunks; ++ch) {
            float* wcur = wbuf + (ch & 1) * kWFloats;
            if (ch + 1 < nchunks) {
                load_chunk(Wg, K, N, WS, (ch + 1) * kKC, wbuf + ((ch + 1) & 1) * kWFloats);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const int k0 = ch * kKC;
            const int kend = min(kKC, K4 - k0);
            for (int kk = 0; kk < kend; kk += 4) {
                float4 a[RM];
#pragma unroll
                for (int r = 0; r < RM; ++r)
                    a[r] = *reinterpret_cast<const float4*>(in + (rg * RM + r) * AS + k0 + kk);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const int q = cg + 16 * j;
                        if (q < nquads) {
                            const float4 w =
                                *reinterpret_cast<const float4*>(wcur + (kk + t) * WS + 4 * q);
#pragma unroll
                            for (int r = 0; r < RM; ++r) {
                                const float av = t == 0 ? a[r].x : t == 1 ? a[r].y
                                               : t == 2 ? a[r].z : a[r].w;
                                acc[r][j][0] = fmaf(av, w.x, acc[r][j][0]);
                                acc[r][j][1] = fmaf(av, w.y, acc[r][j][1]);
                                acc[r][j][2] = fmaf(av, w.z, acc[r][j][2]);
                                acc[r][j][3] = fmaf(av, w.w, acc[r][j][3]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        // epilogue: bias, optional pre-activation save, activation, write the next layer's input
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const int q = cg + 16 * j;
            if (q < nquads) {
                float bias[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) bias[c] = (4 * q + c < N) ? __ldg(bg + 4 * q + c) : 0.f;
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    const int row = rg * RM + r;
                    float v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float z = acc[r][j][c] + bias[c];
                        if (zsave != nullptr && row < nrows && 4 * q + c < N)
                            zsave[(size_t)row * zstride + 4 * q + c] = z;
                        z = act < 0 ? z : apply_act(z, act);
                        v[c] = (4 * q + c < N) ? z : 0.f;
                    }
                    *reinterpret_cast<float4*>(out + row * AS + 4 * q) =
                        make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        __syncthreads();
    }

    // Whole stack for member e.  buf0 holds the input tile (cols >= sizes[0] up to the next
    // multiple of 4 zeroed).  Returns the buffer holding the last layer's output.
    // zsave (optional): array of L pointers to global pre-activation stores for this tile's
    // first row (row stride = that layer's width); entries may be null.
    __device__ static float* forward(const mb_mlp_desc& d, const float* __restrict__ blob, int e,
                                     float* buf0, float* buf1, float* wbuf, float* const* zsave,
                                     int nrows) {
        float* in = buf0;
        float* out = buf1;
        for (int l = 0; l < d.num_layers; ++l) {
            const int K = d.sizes[l], N = d.sizes[l + 1];
            const float* Wg = blob + w_offset(d, l) + (size_t)e * K * N;
            const float* bg = blob + b_offset(d, l) + (size_t)e * N;
            const int act = (l + 1 < d.num_layers) ? d.activation : -1;
            float* zs = zsave ? zsave[l] : nullptr;
            layer(Wg, bg, K, N, act, in, out, wbuf, zs, N, nrows);
            float* t = in; in = out; out = t;
        }
        return in;
    }
};

using SimtTile64 = SimtTile<4, 4>;   // 64-row tiles, widths <= 256 (PE model, REDQ critics)
using SimtTile32 = SimtTile<2, 8>;   // 32-row tiles, widths <= 512 (TQC critics)

}  // namespace mb
