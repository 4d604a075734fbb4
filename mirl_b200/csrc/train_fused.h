// Launch interface of the persistent fused training kernel (train_fused.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mirl_b200.h"

namespace mb {

constexpr int kFusedMaxG = 21;       // CTAs per ensemble member (7 members x 21 = 147 of 148 SMs)
constexpr int kFusedMaxJobs = 4;     // phase-B jobs per CTA
constexpr int kFusedPartStride = 4 + 2 * 64;   // per row block: [loss, -, -, -, d max_logstd 64, d min_logstd 64]

enum { FUSED_HEAD_NLL = 0, FUSED_HEAD_MSE = 1 };

struct FusedLayer {
    int K, N;            // fan-in, fan-out
    int Kp;              // rows of the augmented weight pack: roundup16(K + 1), row K = bias
    int Ns;              // bf16 row stride of pack rows and of g_l rows (>= roundup16(N), = 8 mod 16)
    int As;              // bf16 row stride of this layer's INPUT activations h_l (>= K + 1, = 8 mod 16)
    int nt_out;          // ceil(N / 8)
    int nt_in;           // ceil(K / 8)
    int ksteps_out;      // roundup16(N) / 16
    float wd;            // weight decay coefficient (mlp.py:170-178)
    int64_t w_off, b_off;    // float offsets of W_l / b_l of member 0 in the parameter blob
    int64_t pack_off;        // bf16 offset of member 0's pack [hi: Kp x Ns][lo: Kp x Ns]
    int64_t h_off;           // bf16 offset of h_l: [hi|lo][E][Bp][As]
    int64_t g_off;           // bf16 offset of g_l: [hi|lo][E][Bp][Ns]
};
struct FusedJob { short layer, strip0, nstrips, pad; };

struct FusedArgs {
    // ---- filled by fused_plan
    int L, E, ne, G, act, Bp, nblk_max;
    int as_max, z_stride, head_stride;
    int stage_bytes, act_bytes, z_bytes, head_bytes;
    FusedLayer ly[MB_MAX_LAYERS];
    int njobs[kFusedMaxG];
    FusedJob jobs[kFusedMaxG][kFusedMaxJobs];
    // ---- filled by the caller
    int e0, O, A, D, head_kind, E_total, per_member_rows;
    int nsteps;
    int64_t first_step, train_length, batch_size, steps_per_epoch, index_stride;
    int64_t maxls_off, minls_off;
    const float* obs; const float* actn; const float* delta; const float* reward; const float* done;
    const float* target;           // MSE head: [B][N] shared by the members
    float clip_error;              // MSE head: > 0 clamps (target - q) to +-clip_error (ddpg_agent.py:142-146)
    const int64_t* index;          // bootstrap index table [E][index_stride] or nullptr
    const float* norm;             // normaliser statistics or nullptr
    float* params; float* exp_avg; float* exp_avg_sq;
    float* target_params; float tau;      // optional fused soft target update
    double beta1_pow0, beta2_pow0;        // beta^t0, t0 = Adam steps already taken
    mb_train_cfg cfg;
    float* ensemble_loss;          // [nsteps][E]
    float* loss_terms;             // [nsteps][3]
    // ---- workspace pointers, filled by launch_train_fused
    unsigned* barrier; long long* dbg; float* block_part;   // [E][nblk_max][kFusedPartStride]
    __nv_bfloat16* pack; __nv_bfloat16* hbuf; __nv_bfloat16* gbuf;
};
struct FusedSizes { size_t smem, pack_bytes, h_bytes, g_bytes, part_bytes; };

bool fused_train_supported(const mb_mlp_desc& d, int ne);
size_t fused_train_workspace_bytes(const mb_mlp_desc& d, int64_t batch_size);
int fused_plan(const mb_mlp_desc& d, int ne, int64_t batch_size, int nsm, FusedArgs& a, FusedSizes& sz);
int launch_train_fused(const mb_mlp_desc& d, FusedArgs& a, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace mb
