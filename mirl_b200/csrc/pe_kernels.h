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

// Rows grouped by drawn member.  Member e owns the segment perm[e * seg .. e * seg + totals[e])
// (seg = batch size: a segment can never overflow), filled by ONE kernel: every block counts its
// rows per member, reserves a range of each segment with one atomicAdd per member, and scatters.
// Tiles are not materialised: a consumer CTA derives tile t -> (member, first row, count) from the
// E totals (TileIndex below).  No host sync anywhere.
constexpr int kMaxMembers = 32;
struct BucketPlan {
    int* perm;
    int* totals;        // [kMaxMembers] rows per member (device)
    int64_t seg;        // segment stride in perm
    int max_tiles;      // upper bound of the tile count (grid sizing)
};

#ifdef __CUDACC__
struct TileIndex {
    int toff[kMaxMembers + 1];      // first tile of member e; toff[E] = number of tiles
    int total[kMaxMembers];
    int roff[kMaxMembers];          // rows of members < e: position of member e in bucket order
};
// Called by one full warp; the caller synchronises the CTA afterwards.
__device__ __forceinline__ void tile_index_build(TileIndex& ti, const int* __restrict__ totals, int E, int TM) {
    const int lane = threadIdx.x & 31;
    const int tot = lane < E ? totals[lane] : 0;
    int incl = (tot + TM - 1) / TM;
    const int mine = incl;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    int rincl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, rincl, d);
        if (lane >= d) rincl += up;
    }
    ti.total[lane] = tot;
    ti.toff[lane] = incl - mine;
    ti.roff[lane] = rincl - tot;
    if (lane == 31) ti.toff[32] = incl;
}
// tile t -> (x = member, y = first row inside the member's segment, z = rows in the tile)
__device__ __forceinline__ int4 tile_at(const TileIndex& ti, int TM, int t) {
    int e = 0;
    while (e + 1 < kMaxMembers && t >= ti.toff[e + 1]) ++e;
    const int k = t - ti.toff[e];
    const int left = ti.total[e] - k * TM;
    return make_int4(e, k * TM, left < TM ? left : TM, 0);
}
#endif

// Optional fused pool-row output of the rollout step: row b of the batch is written as
// [obs | act | next_obs | reward | done] (2O+A+2 floats, mbpo_collector.py:69-83 field order) to
// row (row_offset + b) of every destination -- the local pool and/or peer-mapped pools of the
// other GPUs (stores over NVLink straight from the kernel, no collective call).
constexpr int kMaxRowDest = 8;
constexpr int kRowPadMax = 8;       // destination row stride <= 2O+A+2 + kRowPadMax floats
struct RowsOut {
    float* dst[kMaxRowDest];
    int n;                    // 0 = disabled
    int64_t row_offset;
    int64_t stride;           // floats per destination row (>= 2O+A+2; columns beyond are zero-filled)
    int ordered;              // 1: batch row b -> row_offset + b.  0: bucket order (rows grouped by
                              //    drawn member, order inside a group unspecified) -- whole tiles are
                              //    then contiguous and leave as one bulk copy per destination
};

// All-member forward mode of the tcgen05 rollout kernel: instead of the bucket plan the tiles walk
// rows [0, B) of every listed member in order, and the head writes each row's
// [mean D | clamped log-std D] (normalised space) as one 2D-float row of out[slot][B][2D].
struct TcForward {
    int mode;                 // 0 = rollout step, 1 = all-member forward
    float* out;               // [n_members][B][2D]
    int64_t B;
    int64_t in_member_rows;   // 0: inputs shared by all members; B: inputs are [E][B][.]
    int total[kMaxMembers];   // B for listed members, 0 otherwise
    int slot_of[kMaxMembers]; // output slot of member e
    int raw_head;             // head row holds the RAW log-std instead of the clamped one (training)
};
// training forward (separate kernel instantiation): pre-activations of hidden layer l, which the
// backward pass needs, are saved to z[l][member][B][H]
struct TcZSave { float* z[MB_MAX_LAYERS]; };

// np.random.choice(elite, B) on the device, bit-exact with NumPy's legacy MT19937 stream (member_draw.cu)
size_t draw_workspace_bytes(int64_t B, int n);
int launch_draw_members(uint32_t* state, uint32_t* snap, const uint8_t* elites, int n, int64_t B, uint8_t* out,
                        void* ws, cudaStream_t s);

size_t bucket_workspace_bytes(int64_t B, int E);
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
// all-member forward on the tensor cores: out[slot][B][2D] = [mean | clamped log-std] of members[slot]
int launch_pe_forward_tc(const PeDev& m, const void* pack, const float* obs, const float* act, int per_member,
                         int64_t B, const int* members, int n_members, float* out, cudaStream_t s);
bool tc_rollout_supported(const mb_mlp_desc& d);
int launch_pe_dist_finish(const PeDev& m, const int32_t* elites, int n_elite, const float* fwd, const float* obs,
                          const uint8_t* idx, const float* noise, int64_t B, int done_kind, float* next_obs,
                          float* reward, float* done, float* delta_mean, float* delta_std, float* next_obs_mean,
                          float* reward_mean, float* reward_std, int penalty_kind, float penalty_coef,
                          float* penalty, float* penalized_reward, cudaStream_t s);

// ---- model training (pe_train.cu)
size_t train_workspace_bytes(const mb_mlp_desc& d, int64_t B);
int launch_pe_train(const PeDev& m, void* tc_pack, const mb_train_cfg& cfg, float* params, float* exp_avg,
                    float* exp_avg_sq, int64_t step, int32_t* step_counter, const float* obs, const float* act,
                    const float* delta, const float* reward, const float* done, int64_t B,
                    int e0, int e1, float* grads_out, float* ensemble_loss, float* loss_terms,
                    void* ws, cudaStream_t s);
// training forward of members [e0, e0+ne) in ONE launch of the tcgen05 tile kernel: re-packs the
// (just updated) weights, saves every hidden layer's pre-activations and the raw head output
int launch_pe_train_fwd_tc(const PeDev& m, void* pack, const float* obs, const float* act, int64_t B, int e0,
                           int ne, float* const* zsave, float* head_raw, cudaStream_t s);

// critic update as a chain of launches (tensor-core layer GEMMs + head + Adam + soft target update):
// MSE head for DDPG/SAC/REDQ critics wider than the persistent kernel covers, quantile-Huber head for TQC
enum { CRITIC_HEAD_MSE = 0, CRITIC_HEAD_QUANTILE = 1 };
size_t critic_chain_workspace_bytes(const mb_mlp_desc& d, int64_t B);
int launch_critic_train_chain(const mb_mlp_desc& d, int head_kind, float* params, float* exp_avg, float* exp_avg_sq,
                              int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                              const float* obs, const float* act, int O, int A, const float* target, int T,
                              int64_t B, float clip_error, float* target_params, float tau, float* ensemble_loss,
                              float* loss_terms, void* ws, cudaStream_t s);

// actor + entropy-coefficient update (sac_agent.py:88-118,170-195) as a chain of launches
size_t actor_workspace_bytes(const mb_mlp_desc& pd, const mb_mlp_desc& cd, int64_t B);
int launch_actor_update(const mb_mlp_desc& pd, float* pparams, float* exp_avg, float* exp_avg_sq,
                        int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                        const mb_policy_cfg& cfg, const mb_mlp_desc& cd, const float* cparams, int reduce_mode,
                        const int32_t* members, int n_members, const float* obs, const float* noise, int O, int A,
                        int64_t B, float* log_alpha, float* la_m1, float* la_m2, float target_entropy,
                        float alpha_fixed, float* info, void* ws, cudaStream_t s);

// ---- row-parallel forward of one small MLP, the whole network in one launch (mlp_rows.cu): the policy at rollout batch sizes
bool mlp_rows_supported(const mb_mlp_desc& d);
size_t mlp_rows_pack_bytes(const mb_mlp_desc& d);
int launch_mlp_rows_forward(const mb_mlp_desc& d, const float* params, int member, const float* x, int64_t B,
                            float* out, void* pack, cudaStream_t s);

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

// ---- policy action (critic.cu): MLP forward (E = 1) + tanh-Gaussian head
int launch_policy_action(const mb_mlp_desc& d, const float* params, const float* obs, int O, int64_t B,
                         const float* eps, const mb_policy_cfg& cfg, float* action, float* log_prob, float* mean,
                         float* std, void* ws, size_t ws_bytes, cudaStream_t s);

}  // namespace mb
