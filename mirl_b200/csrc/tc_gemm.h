// Per-layer ensemble GEMMs on the tcgen05 tensor cores (tc_gemm.cu): the GEMM-shaped work of the
// model-train step (forward, dX, dW of MyEnsembleLinear, mirl/torch_modules/linear.py:143-149 and
// its autograd) and of the critic forward passes (EnsembleQValue / TQCQValue), at batch sizes
// where one member's layer is a single 128-row tile or two.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mirl_b200.h"

namespace mb {

enum { TG_FWD = 0, TG_DX = 1, TG_DW = 2 };

// D[M x N] = A'[M x R] * B'[R x N] per member, R = reduction length.  Pointers are those of slot /
// member 0; *_ss strides advance per activation slot, *_ms per parameter member.
struct TgOp {
    int mode;
    int M, N, R;
    const float* a;        // trans 0: A'[m][r] = a[m*lda + r];  1: A'[m][r] = a[r*lda + m]
    int64_t lda, a_ss;
    int a_trans, a_act;    // a_act: activation applied while loading (-1 none)
    int a_ones_row;        // row of A' that is all ones (DW: bias gradient), -1 none
    const float* b;        // trans 0: B'[r][n] = b[r*ldb + n];  1: B'[r][n] = b[n*ldb + r]
    int64_t ldb, b_ss, b_ms;   // exactly one of b_ss / b_ms is non-zero
    int b_trans;
    const float* bias;     // FWD: [N] per member
    int64_t bias_ms;
    int out_act;           // FWD (-1 none)
    const float* z;        // DX: out *= act'(z[m*ldz + n]); nullptr: no factor (gradient with respect to a network input)
    int64_t ldz, z_ss;
    int z_act;             // DX: MB_ACT_SWISH (0, the default) or MB_ACT_RELU
    const float* w;        // DW: out += wd * w[m*N + n]
    int64_t w_ms;
    float wd;
    float* out;            // out[m*ldo + n]
    int64_t ldo, out_ss, out_ms;   // exactly one of out_ss / out_ms is non-zero (0/0: shared)
    float* out_ones;       // DW: row a_ones_row goes here (dbias [N] per member)
    int64_t out_ones_ms;
    float* l2_acc;         // DW: += 0.5 * wd * sum w^2 (weight-decay loss term)
};

constexpr int kTgMaxOps = 8;
struct TgBatch {
    int nops, nmem, slot0;
    int member[32];        // parameter member of blockIdx.z % nmem; activation slot = slot0 + that index
    TgOp op[kTgMaxOps];
};

int launch_tc_gemm(const TgBatch& b, cudaStream_t s);

}  // namespace mb
