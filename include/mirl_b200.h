/* mirl_b200 -- C ABI of the B200-native probabilistic-ensemble hot path.
 *
 * Drop-in boundary for ONE path of QiZhou1997/MIRL: the ensemble MLP behind
 *   PEModel.step / PEModel.compute_loss        (mirl/agents/mbpo/pe_model.py:252-307)
 *   PEModelTrainer.train_from_torch_batch      (mirl/agents/mbpo/pe_model_trainer.py:85-106)
 *   EnsembleQValue.value / TQCQValue.value     (mirl/values/ensemble_value.py:65-107,
 *                                               mirl/values/distributional_value.py:59-93)
 * The reference has no FFI of its own (pure Python); these are the entry points its Python
 * classes bind through ctypes.  Rules of the ABI:
 *   - extern "C", plain pointers and sizes; no torch / C++ types;
 *   - every data pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - all tensors are fp32, row-major, last dimension contiguous (the reference's layout);
 *   - no internal allocation, no device synchronisation: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), outputs and workspace are caller-owned;
 *   - returns 0 on success, a negative MB_E* code otherwise; mb_last_error() has the text;
 *   - not thread-safe per handle/argument set (the reference is single-threaded).
 */
#ifndef MIRL_B200_H
#define MIRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 1
#define MB_MAX_LAYERS 8

enum { MB_OK = 0, MB_EINVAL = -1, MB_ECUDA = -2, MB_EWORKSPACE = -3, MB_EUNSUPPORTED = -4 };

/* mirl/torch_modules/utils.py:138-149 get_activation (the two the path's configs use) */
enum { MB_ACT_SWISH = 0, MB_ACT_RELU = 1 };

/* mirl/environments/reward_done_functions/<name>.py done_function */
enum {
    MB_DONE_CHEETAH = 0, MB_DONE_HOPPER = 1, MB_DONE_WALKER2D = 2, MB_DONE_MB_ANT = 3,
    MB_DONE_MB_HUMANOID = 4, MB_DONE_INVERTED_PENDULUM = 5
};

/* mirl/agents/mbpo/mopo_collector.py:45-52 penalty_type */
enum { MB_PENALTY_NONE = 0, MB_PENALTY_TOTAL_VAR = 1, MB_PENALTY_MAX_VAR = 2 };

/* mirl/values/ensemble_value.py:91-96 mode */
enum { MB_REDUCE_MIN = 0, MB_REDUCE_MEAN = 1, MB_REDUCE_MAX = 2 };

/* kernel family for the rollout step: hand-written SIMT fp32 FFMA, or tcgen05 (bf16x3 split
 * operands, fp32 TMEM accumulators).  Both are sm_100a CUDA; there is no CPU path. */
enum { MB_KERNEL_SIMT = 0, MB_KERNEL_TCGEN05 = 1 };

/* Shape of an ensemble MLP (mirl/torch_modules/mlp.py:63-99, linear.py:114-133).
 * Parameters live in ONE packed fp32 blob; mb_mlp_*_offset give float offsets:
 *   for l in [0,L):  W_l [E][sizes[l]][sizes[l+1]]   then   b_l [E][sizes[l+1]]
 *   then (probabilistic models only)  max_logstd [E][D], min_logstd [E][D], D = sizes[L]/2
 * every segment starts on a multiple of 4 floats. */
typedef struct {
    int32_t ensemble_size;              /* E */
    int32_t num_layers;                 /* L linear layers, 1..MB_MAX_LAYERS */
    int32_t sizes[MB_MAX_LAYERS + 1];   /* in, hidden..., out */
    int32_t activation;                 /* MB_ACT_*; applied after every layer but the last */
} mb_mlp_desc;

int64_t mb_mlp_weight_offset(const mb_mlp_desc* d, int layer);
int64_t mb_mlp_bias_offset(const mb_mlp_desc* d, int layer);
int64_t mb_mlp_logstd_offset(const mb_mlp_desc* d, int which /*0 max, 1 min*/);
int64_t mb_mlp_param_count(const mb_mlp_desc* d, int probabilistic);

/* A probabilistic-ensemble dynamics model (PEModel, pe_model.py:191-232).
 * norm = [obs_mean O | obs_std O | act_mean A | act_std A | delta_mean O | delta_std O]
 * (Normalizer mean/std, mirl/processors/normalizer.py:6-28; epsilon 1e-6 is applied inside). */
typedef struct {
    mb_mlp_desc mlp;          /* sizes[0] = O + A, sizes[L] = 2*(O+1) */
    int32_t obs_dim;          /* O */
    int32_t act_dim;          /* A */
    const float* params;      /* packed blob, mb_mlp_param_count(&mlp, 1) floats */
    const float* norm;        /* 4*O + 2*A floats */
    const void* tc_pack;      /* mb_pe_tc_pack output, or NULL (required by MB_KERNEL_TCGEN05) */
} mb_pe_model;

int mb_abi_version(void);
const char* mb_last_error(void);

/* ---- (1) rollout step: PEModel.step (pe_model.py:252-283 over base_model.py:54-76) ------
 * member_idx[B] (uint8) = np.random.choice(elite_indices, B) drawn by the caller exactly like
 * pe_model.py:169; noise[B][O+1] = the torch.randn_like draw of pe_model.py:176.
 * Writes next_obs[B][O], reward[B][1], done[B][1] (float 0/1). */
size_t mb_pe_rollout_workspace_bytes(const mb_pe_model* m, int64_t B);
int mb_pe_rollout_step(const mb_pe_model* m, int kernel, const float* obs, const float* act,
                       const uint8_t* member_idx, const float* noise, int64_t B, int done_kind,
                       float* next_obs, float* reward, float* done,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Random-elite selection of PEModel.step, `np.random.choice(elite, B)` (pe_model.py:165-169), on
 * the device and bit-exact with NumPy's legacy global RandomState: the MT19937 stream with
 * randint's masked rejection.  mt_state is device-accessible uint32[627] -- device memory, or
 * pinned host memory, which keeps the state off the copy engines (no queueing behind the caller's
 * bulk transfers): [0,624) key and [624] pos as returned by np.random.get_state(); both are advanced in place to where NumPy would be after the
 * draws (write them back with np.random.set_state to keep host code reproducible), and
 * [625],[626] receive the low/high word of the number of members produced.  That number is B
 * unless the word budget of the workspace (expected attempts + 12 sigma) ran out -- then the
 * caller draws the remaining B - produced members with a second call from the advanced state.
 * elites_host: the n_elites (1..64) member ids in rank order (host memory).  member_idx: device
 * u8[B].  workspace: device scratch of mb_pe_draw_members_workspace_bytes(B, n_elites). */
size_t mb_pe_draw_members_workspace_bytes(int64_t B, int n_elites);
int mb_pe_draw_members(uint32_t* mt_state, const uint8_t* elites_host, int n_elites, int64_t B,
                       uint8_t* member_idx, void* workspace, size_t workspace_bytes, void* stream);

/* The same draw for a CHAIN of calls queued ahead of the host: mt_state is advanced in place (so
 * the next queued call continues the stream without a host round trip) and state_snapshot
 * (device-accessible uint32[627], may be NULL) additionally receives this call's final block, which
 * the host reads when it consumes this draw.  B == 0 writes nothing. */
int mb_pe_draw_members_chained(uint32_t* mt_state, uint32_t* state_snapshot, const uint8_t* elites_host,
                               int n_elites, int64_t B, uint8_t* member_idx, void* workspace,
                               size_t workspace_bytes, void* stream);

/* The same step fused with the collector's pool-row assembly (MBPOCollector.add_sample,
 * mbpo_collector.py:69-83): every transition is written as one row
 * [obs | act | next_obs | reward | done | zero padding] of row_stride floats
 * (2*O + A + 2 <= row_stride <= 2*O + A + 10) into EVERY destination in dst_host[n_dst] (host array
 * of device pointers; 1..8).  Destinations may be peer-mapped buffers of other GPUs: the kernel
 * then pushes the generated transitions over NVLink itself, which replaces the all-gather of
 * transitions into a replicated model pool.
 *   ordered = 1: batch row b lands in row (row_offset + b) (warp-wide stores per row).
 *   ordered = 0: rows land in bucket order -- grouped by drawn member, order inside a group
 *                unspecified, rows [row_offset, row_offset + B) all filled.  A pool is a bag of
 *                samples, so this is the mode the collectors use: a tile of 128 rows is then
 *                contiguous and leaves as ONE bulk copy per destination.  Needs row_stride % 4 == 0
 *                and 16-byte aligned destinations.
 * obs / act may be strided row views (e.g. the next_obs columns of the previous step's rows):
 * *_stride = floats per row. */
int mb_pe_rollout_step_rows(const mb_pe_model* m, int kernel, const float* obs, int64_t obs_stride,
                            const float* act, int64_t act_stride, const uint8_t* member_idx,
                            const float* noise, int64_t B, int done_kind, float* const* dst_host,
                            int n_dst, int64_t row_stride, int64_t row_offset, int ordered,
                            void* workspace, size_t workspace_bytes, void* stream);

/* PEModel.step(return_distribution=True) + MOPOCollector penalty (mopo_collector.py:42-53).
 * elites_host[n_elite]: elite member ids in RANK order (pe_model.py:234-237).
 * info tensors are [n_elite][B][.] in that order; any of them may be NULL.
 * penalized_reward (optional) = reward - penalty_coef * penalty. */
int mb_pe_rollout_step_dist(const mb_pe_model* m, const float* obs, const float* act,
                            const uint8_t* member_idx, const float* noise, int64_t B,
                            int done_kind, const int32_t* elites_host, int n_elite,
                            float* next_obs, float* reward, float* done,
                            float* delta_mean, float* delta_std, float* next_obs_mean,
                            float* reward_mean, float* reward_std,
                            int penalty_kind, float penalty_coef, float* penalty,
                            float* penalized_reward, void* stream);

/* The same step with the all-elite forward on the tensor cores (tcgen05 rollout kernel in forward
 * mode, then one elementwise kernel for the outputs above).  Needs m->tc_pack and a workspace of
 * mb_pe_dist_workspace_bytes(m, B, n_elite) bytes ([n_elite][B][2(O+1)] floats). */
size_t mb_pe_dist_workspace_bytes(const mb_pe_model* m, int64_t B, int n_elite);
int mb_pe_rollout_step_dist_tc(const mb_pe_model* m, const float* obs, const float* act,
                               const uint8_t* member_idx, const float* noise, int64_t B, int done_kind,
                               const int32_t* elites_host, int n_elite, float* next_obs, float* reward,
                               float* done, float* delta_mean, float* delta_std, float* next_obs_mean,
                               float* reward_mean, float* reward_std, int penalty_kind,
                               float penalty_coef, float* penalty, float* penalized_reward,
                               void* workspace, size_t workspace_bytes, void* stream);

/* All-member forward on the tensor cores (normalised inputs): out[k][b] = [mean (O+1) | clamped
 * log-std (O+1)] of member members_host[k] -- the quantities of mb_pe_forward, interleaved per
 * row.  per_member != 0: obs / act are [E][B][.] (row block e belongs to member e). */
int mb_pe_forward_tc(const mb_pe_model* m, const float* obs, const float* act, int per_member, int64_t B,
                     const int32_t* members_host, int n_members, float* out, void* stream);

/* ---- ensemble forward: per-member Gaussian mean / log-std (pe_model.py:110-151) -----------
 * x is obs/act already split: obs[.][O], act[.][A]; per_member=0 -> inputs [B][.] shared by
 * all members, 1 -> [E][B][.].  normalize!=0 applies the obs/action Normalizers.
 * mean, logstd: [E][B][O+1] (logstd after the soft max/min clamp). */
int mb_pe_forward(const mb_pe_model* m, const float* obs, const float* act, int per_member,
                  int normalize, int64_t B, float* mean, float* logstd, void* stream);

/* ---- (2) model loss / training step -------------------------------------------------------
 * PEModel.compute_loss (pe_model.py:285-307 -> :71-108): ensemble_loss[E] and loss[1] on a
 * per-member [E][B][.] or shared [B][.] batch; no gradients (eval_with_dataset,
 * pe_model_trainer.py:108-121). */
typedef struct {
    float weight_decay[MB_MAX_LAYERS];  /* per-layer L2 coefficients (mlp.py:170-178) */
    float bound_reg_coef;               /* pe_model.py:34,81 */
    float reward_coef;                  /* pe_model.py:33,93 */
    int32_t ignore_terminal_state;      /* pe_model.py:85-89 */
    float lr, beta1, beta2, eps;        /* torch.optim.Adam, pe_model_trainer.py:47-51 */
} mb_train_cfg;

int mb_pe_loss(const mb_pe_model* m, const mb_train_cfg* cfg, const float* obs,
               const float* act, const float* delta, const float* reward, const float* done,
               int per_member, int64_t B, float* ensemble_loss, float* loss, void* stream);

/* One PEModelTrainer.train_from_torch_batch (pe_model_trainer.py:85-106): forward, Gaussian
 * NLL + log-std bound + L2, backward, Adam -- for the members [member_begin, member_end) only
 * (member-sharded training: each rank passes its own slice; the others are untouched; an empty slice is a
 * no-op).
 * params/exp_avg/exp_avg_sq: packed blobs (same layout), updated in place.
 * step = 1-based Adam step count AFTER this update; if step_counter (device int32, optional) is
 * given, the call increments it on the device and uses it instead (the launch sequence then has
 * no host-side state, so it can be captured once in a CUDA graph and replayed every step).
 * Batch tensors are [E][B][.].
 * ensemble_loss[E] (entries of this slice written), loss_terms[3] += {sum_e loss_e, l2, bound}
 * over this slice (zeroed by the call).
 * m->tc_pack, when non-NULL (mb_pe_tc_pack_bytes bytes), is used as SCRATCH: the call re-packs the
 * current weights into it and runs the whole forward as one launch of the tcgen05 tile kernel
 * (pre-activations saved for the backward pass); with NULL the forward is one GEMM launch per
 * layer.  A rollout pack must be rebuilt after training either way (the weights changed).
 * Without step_counter the step runs as ONE persistent cooperative launch (see mb_pe_train_steps);
 * MB_TRAIN_FUSED=0 selects the chain of per-layer launches instead. */
size_t mb_pe_train_workspace_bytes(const mb_pe_model* m, int64_t B);
int mb_pe_train_step(const mb_pe_model* m, const mb_train_cfg* cfg, float* params,
                     float* exp_avg, float* exp_avg_sq, int64_t step, int32_t* step_counter,
                     const float* obs, const float* act, const float* delta,
                     const float* reward, const float* done, int64_t B,
                     int member_begin, int member_end,
                     float* ensemble_loss, float* loss_terms,
                     void* workspace, size_t workspace_bytes, void* stream);

/* `nsteps` consecutive training steps in ONE persistent launch, with the bootstrap batches gathered on
 * the device (PEModelTrainer.sufficiently_train's inner loop, pe_model_trainer.py:240,268-272):
 *   dataset tensors obs/act/delta/reward/done are [N][.] rows; index [E][index_stride] (int64) is the
 *   bootstrap table `ensemble_index` (pe_model_trainer.py:179); step t of the schedule
 *   (t = first_step + s, s < nsteps) uses k = t mod ceil(train_length / batch_size),
 *   rows index[e][k*batch_size .. min((k+1)*batch_size, train_length)) for member e -- the reference's
 *   `for j in range(0, train_length, batch_size)` loop, ragged last batch included.
 *   index == NULL: the tensors are one per-member batch [E][batch_size][.] (train_length = batch_size).
 * adam_steps_done = Adam steps already applied to this optimiser state (bias correction continues from
 * there).  ensemble_loss [nsteps][E] and loss_terms [nsteps][3] receive every step's values (zeroed by
 * the call).  An empty member slice is a no-op.  Layer widths up to 256. */
size_t mb_pe_train_steps_workspace_bytes(const mb_pe_model* m, int64_t batch_size);
int mb_pe_train_steps(const mb_pe_model* m, const mb_train_cfg* cfg, float* params, float* exp_avg,
                      float* exp_avg_sq, int64_t adam_steps_done, const float* obs, const float* act,
                      const float* delta, const float* reward, const float* done, const int64_t* index,
                      int64_t index_stride, int64_t train_length, int64_t batch_size, int64_t first_step,
                      int nsteps, int member_begin, int member_end, float* ensemble_loss, float* loss_terms,
                      void* workspace, size_t workspace_bytes, void* stream);

/* gradients only (same maths as the training step, no Adam); grads is a packed blob.
 * Exposed for parity tests of loss.backward(). */
int mb_pe_loss_grads(const mb_pe_model* m, const mb_train_cfg* cfg, const float* obs,
                     const float* act, const float* delta, const float* reward,
                     const float* done, int64_t B, float* grads, float* ensemble_loss,
                     float* loss_terms, void* workspace, size_t workspace_bytes, void* stream);

/* ---- (3) critic-ensemble targets ----------------------------------------------------------
 * Plain ensemble forward [E][B][out] (EnsembleQValue / TQCQValue `only_return_ensemble`). */
/* workspace for the three critic entry points (intermediate activations of the per-layer path
 * used at small batch); workspace may be NULL, which forces the single-launch tile kernels. */
size_t mb_critic_workspace_bytes(const mb_mlp_desc* d, int64_t B);
int mb_critic_forward(const mb_mlp_desc* d, const float* params, const float* obs,
                      const float* act, int obs_dim, int act_dim, int64_t B, float* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* EnsembleQValue.value (ensemble_value.py:65-107) with the subset reduced on the fly, fused with
 * SACAgent.compute_q_target (sac_agent.py:80-85) when td != 0:
 *   batch-wise subset: members_host[n_sub] (np.random.choice(E, n, replace=False), :84-86);
 *   per-row subset:    row_members[n_sub][B] uint8 device (sample_from_ensemble, :10-34);
 *   both NULL:         all members.
 * Only the drawn members are evaluated.  value[B][1];  with td:
 *   value = reward + (1 - terminal) * discount * (reduce - alpha * log_prob). */
int mb_critic_redq_target(const mb_mlp_desc* d, const float* params, const float* obs,
                          const float* act, int obs_dim, int act_dim, int64_t B,
                          const int32_t* members_host, const uint8_t* row_members, int n_sub,
                          int mode, int td, const float* reward, const float* terminal,
                          const float* log_prob, float alpha, float discount, float* value,
                          void* workspace, size_t workspace_bytes, void* stream);

/* TQCQValue.value(return_atoms=True) (distributional_value.py:32-93) fused with
 * TQCAgent.compute_q_target (tqc_agent.py:41-57) when td != 0.  atoms[B][E*Q - n_drop] ascending
 * (unsorted member-major order when n_drop == 0, like the reference); value[B][1] = mean of the
 * kept atoms BEFORE the TD transform (what TQCQValue.value returns). */
int mb_critic_tqc_target(const mb_mlp_desc* d, const float* params, const float* obs,
                         const float* act, int obs_dim, int act_dim, int64_t B, int n_drop,
                         int td, const float* reward, const float* terminal,
                         const float* log_prob, float alpha, float discount,
                         float* value, float* atoms, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- critic update: DDPGAgent.compute_qf_loss (ddpg_agent.py:138-149: F.mse_loss of the [E][B][out]
 * ensemble prediction against the expanded target, optional clip_error) + qf_optimizer.step (torch.optim.Adam,
 * ddpg_agent.py:82-86) + the soft target update ptu.soft_update_from_to (torch_modules/utils.py:151-155), as ONE
 * persistent launch (the kernel of mb_pe_train_steps with an MSE head): forward, loss, backward, Adam and
 * target <- (1 - tau) target + tau params.  obs/act [B][.] are shared by the members; q_target [B][out].
 * target_params (same blob layout) may be NULL (no target update this step, sac_agent.py:165-166).
 * ensemble_loss[E] = each member's share of the loss; loss_terms[0] = the loss.  Layer widths up to 256 run the
 * persistent kernel; wider critics run the same update as a chain of launches (tcgen05 layer GEMMs, head, Adam,
 * soft update). */
size_t mb_critic_train_workspace_bytes(const mb_mlp_desc* d, int64_t B);
int mb_critic_train_step(const mb_mlp_desc* d, float* params, float* exp_avg, float* exp_avg_sq,
                         int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                         const float* obs, const float* act, int obs_dim, int act_dim, const float* q_target,
                         int64_t B, float clip_error, float* target_params, float tau, float* ensemble_loss,
                         float* loss_terms, void* workspace, size_t workspace_bytes, void* stream);

/* TQC critic update: TQCAgent.compute_qf_loss (tqc_agent.py:59-66) = quantile_huber_loss_f (tqc_agent.py:5-14) of
 * the [E][B][Q] ensemble atoms (TQCQValue.value(only_return_ensemble=True), distributional_value.py:73-77) against
 * atoms_target[B][n_target] (mb_critic_tqc_target's atoms), + Adam + soft target update as above.  d->sizes[L] = Q.
 * Workspace: mb_critic_train_workspace_bytes.  configs/tqc.json: E 5, 3 x 512 relu, Q 25, n_target 115. */
int mb_tqc_train_step(const mb_mlp_desc* d, float* params, float* exp_avg, float* exp_avg_sq,
                      int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                      const float* obs, const float* act, int obs_dim, int act_dim, const float* atoms_target,
                      int n_target, int64_t B, float* target_params, float tau, float* ensemble_loss,
                      float* loss_terms, void* workspace, size_t workspace_bytes, void* stream);

/* ---- policy action inside the rollout loop: GaussianPolicy.action -> MeanLogstdGaussian.forward
 * (mirl/policies/gaussian_policy.py:35-47, mirl/torch_modules/gaussian.py:55-97,199-204).  d describes the policy MLP
 * as a 1-member ensemble (sizes[0] = obs_dim, sizes[L] = 2 * act_dim); eps[B][A] is the Normal.rsample() / sample()
 * draw (NULL = deterministic: the mean).  Writes action[B][A] and, when non-NULL, log_prob[B][1], mean / std [B][A]
 * (after mean_scale / std_scale / min_std).  workspace >= mb_policy_workspace_bytes. */
enum { MB_LOGSTD_TANH = 0, MB_LOGSTD_CLAMP = 1, MB_LOGSTD_NO = 2 };   /* gaussian.py:17-30 bound_logstd */
typedef struct {
    float mean_scale, std_scale, min_std, logstd_extra_bias;   /* gaussian.py:34-44,170-197 */
    int32_t bound_mode;                                         /* MB_LOGSTD_* */
    int32_t squashed;                                           /* tanh squashing (gaussian_policy.py:13) */
} mb_policy_cfg;
size_t mb_policy_workspace_bytes(const mb_mlp_desc* d, int64_t B);
int mb_policy_action(const mb_mlp_desc* d, const float* params, const float* obs, int obs_dim, int64_t B,
                     const float* eps, const mb_policy_cfg* cfg, float* action, float* log_prob, float* mean,
                     float* std, void* workspace, size_t workspace_bytes, void* stream);

/* ---- actor + entropy-coefficient update (second half): SACAgent.train_from_torch_batch's "update actor
 * and alpha" block (sac_agent.py:170-195) --
 *   new_action, info = policy.action(obs, reparameterize=True, return_log_prob=True)        (gaussian.py:55-97)
 *   compute_alpha_loss (sac_agent.py:88-100) + alpha_optimizer.step()                       (Adam on log_alpha)
 *   compute_policy_loss (sac_agent.py:102-118): alpha * log_prob.mean() - qf.value(obs, new_action, **v_pi).mean()
 *   policy_optimizer.step()                                                                  (torch.optim.Adam)
 * as a chain of launches (tcgen05 layer GEMMs forward and backward, hand-written heads); the critics are read, not
 * updated.  policy: 1-member MLP obs -> 2 * act_dim; critic: the E-member ensemble on [obs | action] (outputs 1, or
 * Q atoms for TQC).  reduce_mode MB_REDUCE_* over `members` (host array of n_members indices, NULL = all; the
 * batch-wise draw of ensemble_value.py:84-86 -- per-row draws are not covered); with Q > 1 outputs the value is the mean
 * of the listed members' atoms (distributional_value.py:51-55, number_drop 0) and reduce_mode must be MB_REDUCE_MEAN.
 * noise[B][act_dim] = the rsample() draw.  log_alpha / its two Adam moments are device scalars (NULL = fixed
 * alpha_fixed, use_automatic_entropy_tuning False); both optimisers take lr / betas / eps and have made
 * adam_steps_done steps.  info (device float[8], may be NULL): [0] policy/loss [1] policy/q_pi [2] policy/entropy
 * [3] alpha/value [4] alpha/loss [5] the alpha the policy loss used. */
size_t mb_actor_workspace_bytes(const mb_mlp_desc* policy, const mb_mlp_desc* critic, int64_t B);
int mb_actor_update(const mb_mlp_desc* policy, float* policy_params, float* exp_avg, float* exp_avg_sq,
                    int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                    const mb_policy_cfg* cfg, const mb_mlp_desc* critic, const float* critic_params,
                    int reduce_mode, const int32_t* members, int n_members,
                    const float* obs, const float* noise, int obs_dim, int act_dim, int64_t B,
                    float* log_alpha, float* alpha_exp_avg, float* alpha_exp_avg_sq, float target_entropy,
                    float alpha_fixed, float* info, void* workspace, size_t workspace_bytes, void* stream);

/* ---- tcgen05 operand pack: bf16 hi/lo split of every member's weights in the UMMA
 * shared-memory layout the rollout kernel streams (rebuild after the weights change). */
size_t mb_pe_tc_pack_bytes(const mb_pe_model* m);
int mb_pe_tc_pack(const mb_pe_model* m, void* pack, void* stream);

/* ---- peer-mapped pools: the multi-GPU destinations of mb_pe_rollout_step_rows -------------------
 * One process per GPU.  Every rank allocates its replica of the model pool with
 * mb_peer_pool_alloc (cudaMalloc + CUDA IPC export; `handle` receives MB_IPC_HANDLE_BYTES bytes),
 * the ranks exchange the handles over any host channel (torch.distributed all_gather_object in
 * mirl_b200.dist.PeerPools), and mb_peer_pool_open maps a peer's replica into this process
 * (peer access over NVLink is enabled by the driver).  The mapped pointers are then passed as
 * destinations; close every mapping before the owner frees its pool. */
#define MB_IPC_HANDLE_BYTES 64
int mb_peer_pool_alloc(size_t bytes, void** dev_ptr, unsigned char* handle);
int mb_peer_pool_open(const unsigned char* handle, void** dev_ptr);
int mb_peer_pool_close(void* dev_ptr);
int mb_peer_pool_free(void* dev_ptr);

/* Persistent kernels launch one CTA per SM.  When other kernels have to run under them -- the
 * single-CTA MT19937 generator of the next call's member draw (mb_pe_draw_members on a second
 * stream), or the communication kernels of a collective -- leave `n` SMs free for them.
 * 0 (default) = use every SM. */
int mb_set_reserved_sms(int n);

/* ---- instrumentation for bench.py's roofline leg (not used on the product path) ---------------
 * When enabled, mb_pe_rollout_step brackets its dominant kernel (the fused MLP kernel, not the
 * bucketing pre-pass) with CUDA events on `stream`; mb_profile_last_ms waits for the end event
 * and returns the elapsed milliseconds of the last bracketed launch. */
int mb_profile_enable(int on);
int mb_profile_last_ms(float* ms);

#ifdef __cplusplus
}
#endif
#endif /* MIRL_B200_H */
