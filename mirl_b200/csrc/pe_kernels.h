// Internal launcher interface between the C ABI (c_api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mirl_b200.h"

namespace mb {

// Device-side view of mb_pe_model (passed by value as a kernel parameter).
struct PeDev {
    mb_mlp_desc mlp;
    int O, A;
    const float* params;
    const float* norm;
};

// Rows grouped by drawn member: perm[start .. start+count) are the rows of tile t = tiles[t]
// (x = member, y = start, z = count).  *ntiles is computed on the device (no host sync).
struct BucketPlan {
    int* perm;
    int4* tiles;
    int* ntiles;
    int max_tiles;
};
// Optional fused pool-row output of the rollout step: row b of the batch is written as
// [obs | act | next_obs | reward | done] (2O+A+2 floats, mbpo_collector.py:69-83 field order) to
// row (row_offset + b) of every destination -- the local pool and/or peer-mapped pools of the
// other GPUs (stores over NVLink straight from the kernel, no collective call).
constexpr int kMaxRowDest = 8;
struct RowsOut {
    float* dst[kMaxRowDest];
    int n;                    // 0 = disabled
    int64_t row_offset;
};

// np.random.choice(elite, B) on the device, bit-exact with NumPy's legacy MT19937 stream (member_draw.cu)
size_t draw_workspace_bytes(int64_t B, int n);
int launch_draw_members(uint32_t* state, const uint8_t* elites, int n, int64_t B, uint8_t* out, void* ws,
                        cudaStream_t s);

size_t bucket_workspace_bytes(int64_t B, int TM);
BucketPlan bucket_rows(const uint8_t* idx, int64_t B, int E, int TM, void* ws, cudaStream_t s);

int reserved_sms();     // mb_set_reserved_sms (c_api.cu)

// bench instrumentation (c_api.cu): events around the dominant rollout kernel when enabled
void prof_begin(cudaStream_t s);
void prof_end(cudaStream_t s);

// ---- SIMT fp32 family (pe_simt.cu)
int launch_pe_rollout_simt(const PeDev& m, const float* obs, int64_t obs_stride, const float* act,
                           int64_t act_stride, const uint8_t* idx, const float* noise, int64_t B,
                           int done_kind, float* next_obs, float* reward, float* done,
                           const RowsOut& rows, void* ws, cudaStream_t s);
int launch_pe_forward_simt(const PeDev& m, const float* obs, const float* act, int per_member,
                           int normalize, int64_t B, float* mean, float* logstd, const float* delta,
                           const float* reward, const float* done, int ignore_terminal,
                           float reward_coef, float* ensemble_loss, cudaStream_t s);
int launch_pe_rollout_dist_simt(const PeDev& m, const int32_t* elites, int n_elite, const float* obs,
                                const float* act, const uint8_t* idx, const float* noise, int64_t B,
                                int done_kind, float* next_obs, float* reward, float* done,
                                float* delta_mean, float* delta_std, float* next_obs_mean,
                                float* reward_mean, float* reward_std, int penalty_kind,
                                float penalty_coef, float* penalty, float* penalized_reward,
                                cudaStream_t s);

// ---- tcgen05 family (pe_tc.cu)
size_t tc_pack_bytes(const mb_mlp_desc& d);
int launch_tc_pack(const PeDev& m, void* pack, cudaStream_t s);
int launch_pe_rollout_tc(const PeDev& m, const void* pack, const float* obs, int64_t obs_stride,
                         const float* act, int64_t act_stride, const uint8_t* idx, const float* noise,
                         int64_t B, int done_kind, float* next_obs, float* reward, float* done,
                         const RowsOut& rows, void* ws, cudaStream_t s);

// ---- model training (pe_train.cu)
size_t train_workspace_bytes(const mb_mlp_desc& d, int64_t B);
int launch_pe_train(const PeDev& m, const mb_train_cfg& cfg, float* params, float* exp_avg,
                    float* exp_avg_sq, int64_t step, int32_t* step_counter, const float* obs, const float* act,
                    const float* delta, const float* reward, const float* done, int64_t B,
                    int e0, int e1, float* grads_out, float* ensemble_loss, float* loss_terms,
                    void* ws, cudaStream_t s);

// ---- critic targets (critic.cu)
size_t critic_workspace_bytes(const mb_mlp_desc& d, int64_t B);
int launch_critic_forward(const mb_mlp_desc& d, const float* params, const float* obs,
                          const float* act, int O, int A, int64_t B, float* out, void* ws, cudaStream_t s);
int launch_critic_redq(const mb_mlp_desc& d, const float* params, const float* obs, const float* act,
                       int O, int A, int64_t B, const int32_t* members, const uint8_t* row_members,
                       int n_sub, int mode, int td, const float* reward, const float* terminal,
                       const float* log_prob, float alpha, float discount, float* value, void* ws,
                       cudaStream_t s);
int launch_critic_tqc(const mb_mlp_desc& d, const float* params, const float* obs, const float* act,
                      int O, int A, int64_t B, int n_drop, int td, const float* reward,
                      const float* terminal, const float* log_prob, float alpha, float discount,
                      float* value, float* atoms, void* ws, cudaStream_t s);

}  // namespace mb
