// extern "C" surface of libmirl_b200.so (declared in include/mirl_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include <cstring>
#include "pe_kernels.h"
#include "simt_mlp.cuh"
#include "train_fused.h"
#include <math.h>
#include <stdlib.h>

namespace mb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int validate_desc(const mb_mlp_desc* d, const char* who) {
    MB_CHECK_ARG(d != nullptr, "%s: null mlp descriptor", who);
    MB_CHECK_ARG(d->num_layers >= 1 && d->num_layers <= MB_MAX_LAYERS, "%s: num_layers %d outside [1,%d]",
                 who, d->num_layers, MB_MAX_LAYERS);
    MB_CHECK_ARG(d->ensemble_size >= 1 && d->ensemble_size <= 32, "%s: ensemble_size %d outside [1,32]", who,
                 d->ensemble_size);
    for (int l = 0; l <= d->num_layers; ++l)
        MB_CHECK_ARG(d->sizes[l] >= 1 && d->sizes[l] <= SimtTile32::MAXW, "%s: layer width %d outside [1,%d]",
                     who, d->sizes[l], SimtTile32::MAXW);
    MB_CHECK_ARG(d->activation == MB_ACT_SWISH || d->activation == MB_ACT_RELU, "%s: unknown activation %d",
                 who, d->activation);
    return MB_OK;
}

int validate_model(const mb_pe_model* m, const char* who) {
    MB_CHECK_ARG(m != nullptr, "%s: null model", who);
    int rc = validate_desc(&m->mlp, who);
    if (rc) return rc;
    const mb_mlp_desc& d = m->mlp;
    MB_CHECK_ARG(m->obs_dim >= 1 && m->act_dim >= 1, "%s: bad obs/act dims", who);
    MB_CHECK_ARG(d.sizes[0] == m->obs_dim + m->act_dim, "%s: mlp input %d != obs+act %d", who, d.sizes[0],
                 m->obs_dim + m->act_dim);
    MB_CHECK_ARG(d.sizes[d.num_layers] == 2 * (m->obs_dim + 1), "%s: mlp output %d != 2*(obs+1) %d", who,
                 d.sizes[d.num_layers], 2 * (m->obs_dim + 1));
    MB_CHECK_ARG(m->obs_dim + 1 <= 64, "%s: obs_dim %d too large (max 63)", who, m->obs_dim);
    for (int l = 0; l <= d.num_layers; ++l)
        MB_CHECK_ARG(d.sizes[l] <= SimtTile64::MAXW, "%s: model layer width %d > %d", who, d.sizes[l],
                     SimtTile64::MAXW);
    MB_CHECK_ARG(m->params != nullptr && m->norm != nullptr, "%s: null params/norm", who);
    return MB_OK;
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("MB_PDL"); return !(v && v[0] == '0'); }();
    return on;
}

static int g_reserved_sms = 0;
int reserved_sms() { return g_reserved_sms; }

static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[2] = {nullptr, nullptr};
void prof_begin(cudaStream_t s) {
    if (g_prof_on) cudaEventRecord(g_prof_ev[0], s);
}
void prof_end(cudaStream_t s) {
    if (g_prof_on) cudaEventRecord(g_prof_ev[1], s);
}

static PeDev to_dev(const mb_pe_model* m) {
    PeDev d;
    d.mlp = m->mlp;
    d.O = m->obs_dim;
    d.A = m->act_dim;
    d.params = m->params;
    d.norm = m->norm;
    return d;
}

}  // namespace mb

using namespace mb;

extern "C" {

int mb_abi_version(void) { return MB_ABI_VERSION; }
const char* mb_last_error(void) { return g_err; }

int64_t mb_mlp_weight_offset(const mb_mlp_desc* d, int layer) { return w_offset(*d, layer); }
int64_t mb_mlp_bias_offset(const mb_mlp_desc* d, int layer) { return b_offset(*d, layer); }
int64_t mb_mlp_logstd_offset(const mb_mlp_desc* d, int which) { return logstd_offset(*d, which); }
int64_t mb_mlp_param_count(const mb_mlp_desc* d, int probabilistic) {
    return param_count(*d, probabilistic != 0);
}

size_t mb_pe_rollout_workspace_bytes(const mb_pe_model* m, int64_t B) {
    if (m == nullptr) return 0;
    return bucket_workspace_bytes(B < 1 ? 1 : B, m->mlp.ensemble_size);
}

static int rollout_common(const char* who, const mb_pe_model* m, int kernel, const float* obs, int64_t obs_stride,
                          const float* act, int64_t act_stride, const uint8_t* member_idx, const float* noise,
                          int64_t B, int done_kind, float* next_obs, float* reward, float* done,
                          const RowsOut& rows, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_model(m, who);
    if (rc) return rc;
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "%s: B %lld out of range", who, (long long)B);
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(obs && act && member_idx && noise, "%s: null input tensor", who);
    MB_CHECK_ARG(obs_stride >= m->obs_dim && act_stride >= m->act_dim, "%s: row strides smaller than the rows", who);
    MB_CHECK_ARG(done_kind >= MB_DONE_CHEETAH && done_kind <= MB_DONE_INVERTED_PENDULUM,
                 "%s: unknown done_kind %d", who, done_kind);
    if (workspace == nullptr || workspace_bytes < mb_pe_rollout_workspace_bytes(m, B)) {
        set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, mb_pe_rollout_workspace_bytes(m, B));
        return MB_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (kernel == MB_KERNEL_SIMT)
        return launch_pe_rollout_simt(to_dev(m), obs, obs_stride, act, act_stride, member_idx, noise, B, done_kind,
                                      next_obs, reward, done, rows, workspace, s);
    if (kernel == MB_KERNEL_TCGEN05) {
        MB_CHECK_ARG(m->tc_pack != nullptr, "%s: MB_KERNEL_TCGEN05 needs tc_pack", who);
        return launch_pe_rollout_tc(to_dev(m), m->tc_pack, obs, obs_stride, act, act_stride, member_idx, noise, B,
                                    done_kind, next_obs, reward, done, rows, workspace, s);
    }
    set_error("%s: unknown kernel family %d", who, kernel);
    return MB_EINVAL;
}

int mb_pe_rollout_step(const mb_pe_model* m, int kernel, const float* obs, const float* act,
                       const uint8_t* member_idx, const float* noise, int64_t B, int done_kind,
                       float* next_obs, float* reward, float* done, void* workspace,
                       size_t workspace_bytes, void* stream) {
    MB_CHECK_ARG(next_obs && reward && done, "mb_pe_rollout_step: null output tensor");
    RowsOut rows{};
    return rollout_common("mb_pe_rollout_step", m, kernel, obs, m ? m->obs_dim : 0, act, m ? m->act_dim : 0,
                          member_idx, noise, B, done_kind, next_obs, reward, done, rows, workspace,
                          workspace_bytes, stream);
}

int mb_pe_rollout_step_rows(const mb_pe_model* m, int kernel, const float* obs, int64_t obs_stride,
                            const float* act, int64_t act_stride, const uint8_t* member_idx,
                            const float* noise, int64_t B, int done_kind, float* const* dst_host,
                            int n_dst, int64_t row_stride, int64_t row_offset, int ordered,
                            void* workspace, size_t workspace_bytes, void* stream) {
    MB_CHECK_ARG(dst_host && n_dst >= 1 && n_dst <= kMaxRowDest, "mb_pe_rollout_step_rows: n_dst %d outside [1,%d]",
                 n_dst, kMaxRowDest);
    MB_CHECK_ARG(row_offset >= 0, "mb_pe_rollout_step_rows: negative row_offset");
    MB_CHECK_ARG(m != nullptr, "mb_pe_rollout_step_rows: null model");
    const int64_t W = 2 * (int64_t)m->obs_dim + m->act_dim + 2;
    MB_CHECK_ARG(row_stride >= W && row_stride <= W + kRowPadMax,
                 "mb_pe_rollout_step_rows: row_stride %lld outside [%lld,%lld]", (long long)row_stride, (long long)W,
                 (long long)(W + kRowPadMax));
    // bucket order moves whole tiles by 16-byte-granular bulk copies
    MB_CHECK_ARG(ordered || row_stride % 4 == 0, "mb_pe_rollout_step_rows: ordered=0 needs row_stride %% 4 == 0 (got %lld)",
                 (long long)row_stride);
    RowsOut rows{};
    rows.n = n_dst;
    rows.row_offset = row_offset;
    rows.stride = row_stride;
    rows.ordered = ordered ? 1 : 0;
    for (int i = 0; i < n_dst; ++i) {
        MB_CHECK_ARG(dst_host[i] != nullptr, "mb_pe_rollout_step_rows: null destination %d", i);
        MB_CHECK_ARG(ordered || (reinterpret_cast<uintptr_t>(dst_host[i]) & 15) == 0,
                     "mb_pe_rollout_step_rows: ordered=0 needs 16-byte aligned destinations");
        rows.dst[i] = dst_host[i];
    }
    return rollout_common("mb_pe_rollout_step_rows", m, kernel, obs, obs_stride, act, act_stride, member_idx, noise,
                          B, done_kind, nullptr, nullptr, nullptr, rows, workspace, workspace_bytes, stream);
}

int mb_pe_rollout_step_dist(const mb_pe_model* m, const float* obs, const float* act,
                            const uint8_t* member_idx, const float* noise, int64_t B, int done_kind,
                            const int32_t* elites_host, int n_elite, float* next_obs, float* reward,
                            float* done, float* delta_mean, float* delta_std, float* next_obs_mean,
                            float* reward_mean, float* reward_std, int penalty_kind,
                            float penalty_coef, float* penalty, float* penalized_reward,
                            void* stream) {
    int rc = validate_model(m, "mb_pe_rollout_step_dist");
    if (rc) return rc;
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "mb_pe_rollout_step_dist: B out of range");
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(obs && act && member_idx && noise && next_obs && reward && done,
                 "mb_pe_rollout_step_dist: null tensor");
    MB_CHECK_ARG(elites_host && n_elite >= 1 && n_elite <= m->mlp.ensemble_size,
                 "mb_pe_rollout_step_dist: bad elite list");
    for (int i = 0; i < n_elite; ++i)
        MB_CHECK_ARG(elites_host[i] >= 0 && elites_host[i] < m->mlp.ensemble_size,
                     "mb_pe_rollout_step_dist: elite id %d out of range", elites_host[i]);
    MB_CHECK_ARG(penalty_kind >= MB_PENALTY_NONE && penalty_kind <= MB_PENALTY_MAX_VAR,
                 "mb_pe_rollout_step_dist: unknown penalty kind %d", penalty_kind);
    return launch_pe_rollout_dist_simt(to_dev(m), elites_host, n_elite, obs, act, member_idx, noise, B,
                                       done_kind, next_obs, reward, done, delta_mean, delta_std,
                                       next_obs_mean, reward_mean, reward_std, penalty_kind, penalty_coef,
                                       penalty, penalized_reward, (cudaStream_t)stream);
}

size_t mb_pe_dist_workspace_bytes(const mb_pe_model* m, int64_t B, int n_elite) {
    if (m == nullptr || B < 1 || n_elite < 1) return 0;
    return (size_t)n_elite * (size_t)B * 2 * (size_t)(m->obs_dim + 1) * sizeof(float) + 256;
}

int mb_pe_rollout_step_dist_tc(const mb_pe_model* m, const float* obs, const float* act,
                               const uint8_t* member_idx, const float* noise, int64_t B, int done_kind,
                               const int32_t* elites_host, int n_elite, float* next_obs, float* reward,
                               float* done, float* delta_mean, float* delta_std, float* next_obs_mean,
                               float* reward_mean, float* reward_std, int penalty_kind,
                               float penalty_coef, float* penalty, float* penalized_reward,
                               void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_model(m, "mb_pe_rollout_step_dist_tc");
    if (rc) return rc;
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "mb_pe_rollout_step_dist_tc: B out of range");
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(obs && act && member_idx && noise && next_obs && reward && done,
                 "mb_pe_rollout_step_dist_tc: null tensor");
    MB_CHECK_ARG(m->tc_pack != nullptr, "mb_pe_rollout_step_dist_tc: needs tc_pack");
    MB_CHECK_ARG(elites_host && n_elite >= 1 && n_elite <= m->mlp.ensemble_size,
                 "mb_pe_rollout_step_dist_tc: bad elite list");
    for (int i = 0; i < n_elite; ++i)
        MB_CHECK_ARG(elites_host[i] >= 0 && elites_host[i] < m->mlp.ensemble_size,
                     "mb_pe_rollout_step_dist_tc: elite id %d out of range", elites_host[i]);
    MB_CHECK_ARG(penalty_kind >= MB_PENALTY_NONE && penalty_kind <= MB_PENALTY_MAX_VAR,
                 "mb_pe_rollout_step_dist_tc: unknown penalty kind %d", penalty_kind);
    if (workspace == nullptr || workspace_bytes < mb_pe_dist_workspace_bytes(m, B, n_elite)) {
        set_error("mb_pe_rollout_step_dist_tc: workspace %zu < %zu bytes", workspace_bytes,
                  mb_pe_dist_workspace_bytes(m, B, n_elite));
        return MB_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* fwd = reinterpret_cast<float*>(workspace);
    rc = launch_pe_forward_tc(to_dev(m), m->tc_pack, obs, act, 0, B, elites_host, n_elite, fwd, s);
    if (rc) return rc;
    return launch_pe_dist_finish(to_dev(m), elites_host, n_elite, fwd, obs, member_idx, noise, B, done_kind,
                                 next_obs, reward, done, delta_mean, delta_std, next_obs_mean, reward_mean,
                                 reward_std, penalty_kind, penalty_coef, penalty, penalized_reward, s);
}

int mb_pe_forward_tc(const mb_pe_model* m, const float* obs, const float* act, int per_member, int64_t B,
                     const int32_t* members_host, int n_members, float* out, void* stream) {
    int rc = validate_model(m, "mb_pe_forward_tc");
    if (rc) return rc;
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "mb_pe_forward_tc: B out of range");
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(obs && act && out && members_host, "mb_pe_forward_tc: null tensor");
    MB_CHECK_ARG(m->tc_pack != nullptr, "mb_pe_forward_tc: needs tc_pack");
    MB_CHECK_ARG(n_members >= 1 && n_members <= m->mlp.ensemble_size, "mb_pe_forward_tc: bad member count %d", n_members);
    for (int i = 0; i < n_members; ++i)
        MB_CHECK_ARG(members_host[i] >= 0 && members_host[i] < m->mlp.ensemble_size,
                     "mb_pe_forward_tc: member id %d out of range", members_host[i]);
    return launch_pe_forward_tc(to_dev(m), m->tc_pack, obs, act, per_member, B, members_host, n_members, out,
                                (cudaStream_t)stream);
}

int mb_pe_forward(const mb_pe_model* m, const float* obs, const float* act, int per_member,
                  int normalize, int64_t B, float* mean, float* logstd, void* stream) {
    int rc = validate_model(m, "mb_pe_forward");
    if (rc) return rc;
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "mb_pe_forward: B out of range");
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(obs && act && mean && logstd, "mb_pe_forward: null tensor");
    return launch_pe_forward_simt(to_dev(m), obs, act, per_member, normalize, B, mean, logstd, nullptr,
                                  nullptr, nullptr, 0, 1.f, nullptr, (cudaStream_t)stream);
}

int mb_pe_loss(const mb_pe_model* m, const mb_train_cfg* cfg, const float* obs, const float* act,
               const float* delta, const float* reward, const float* done, int per_member, int64_t B,
               float* ensemble_loss, float* loss, void* stream) {
    int rc = validate_model(m, "mb_pe_loss");
    if (rc) return rc;
    MB_CHECK_ARG(cfg != nullptr, "mb_pe_loss: null cfg");
    MB_CHECK_ARG(B >= 1 && B < (int64_t)1 << 31, "mb_pe_loss: B out of range");
    MB_CHECK_ARG(obs && act && delta && reward && ensemble_loss, "mb_pe_loss: null tensor");
    MB_CHECK_ARG(!cfg->ignore_terminal_state || done, "mb_pe_loss: ignore_terminal_state needs done");
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(ensemble_loss, 0, sizeof(float) * m->mlp.ensemble_size, s);
    (void)loss;   // the scalar loss (sum + L2 + bound) is assembled by the host wrapper
    return launch_pe_forward_simt(to_dev(m), obs, act, per_member, 1, B, nullptr, nullptr, delta, reward,
                                  done, cfg->ignore_terminal_state, cfg->reward_coef, ensemble_loss, s);
}

// MB_TRAIN_FUSED=0 keeps the round-1 chain of per-layer launches (A/B timing, gradient-only mode)
static bool fused_enabled() {
    static const bool on = [] { const char* v = getenv("MB_TRAIN_FUSED"); return !(v && v[0] == '0'); }();
    return on;
}

size_t mb_pe_train_workspace_bytes(const mb_pe_model* m, int64_t B) {
    const size_t chain = train_workspace_bytes(m->mlp, B < 1 ? 1 : B);
    const size_t fused = fused_train_supported(m->mlp, 1) ? fused_train_workspace_bytes(m->mlp, B < 1 ? 1 : B) : 0;
    return chain > fused ? chain : fused;
}

size_t mb_pe_train_steps_workspace_bytes(const mb_pe_model* m, int64_t batch_size) {
    if (!m || !fused_train_supported(m->mlp, 1)) return 0;
    return fused_train_workspace_bytes(m->mlp, batch_size < 1 ? 1 : batch_size);
}

// common argument block of the fused kernel for a probabilistic-ensemble model
static void fused_pe_args(FusedArgs& a, const mb_pe_model* m, const mb_train_cfg* cfg, float* params,
                          float* exp_avg, float* exp_avg_sq, int64_t steps_done, int e0, int e1) {
    const mb_mlp_desc& d = m->mlp;
    a.e0 = e0;
    a.ne = e1 - e0;
    a.O = m->obs_dim;
    a.A = m->act_dim;
    a.D = m->obs_dim + 1;
    a.head_kind = FUSED_HEAD_NLL;
    a.E_total = d.ensemble_size;
    a.maxls_off = logstd_offset(d, 0);
    a.minls_off = logstd_offset(d, 1);
    a.norm = m->norm;
    a.params = params;
    a.exp_avg = exp_avg;
    a.exp_avg_sq = exp_avg_sq;
    a.target_params = nullptr;
    a.tau = 0.f;
    a.target = nullptr;
    a.cfg = *cfg;
    a.beta1_pow0 = pow((double)cfg->beta1, (double)steps_done);
    a.beta2_pow0 = pow((double)cfg->beta2, (double)steps_done);
    for (int l = 0; l < d.num_layers; ++l) a.ly[l].wd = cfg->weight_decay[l];
}

int mb_pe_train_steps(const mb_pe_model* m, const mb_train_cfg* cfg, float* params, float* exp_avg,
                      float* exp_avg_sq, int64_t adam_steps_done, const float* obs, const float* act,
                      const float* delta, const float* reward, const float* done, const int64_t* index,
                      int64_t index_stride, int64_t train_length, int64_t batch_size, int64_t first_step,
                      int nsteps, int member_begin, int member_end, float* ensemble_loss, float* loss_terms,
                      void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "mb_pe_train_steps";
    int rc = validate_model(m, who);
    if (rc) return rc;
    MB_CHECK_ARG(cfg != nullptr, "%s: null cfg", who);
    MB_CHECK_ARG(params && exp_avg && exp_avg_sq && params == m->params, "%s: params must be the model's blob", who);
    MB_CHECK_ARG(nsteps >= 0 && nsteps <= (1 << 20), "%s: nsteps %d out of range", who, nsteps);
    MB_CHECK_ARG(batch_size >= 1 && batch_size < (int64_t)1 << 20, "%s: batch_size out of range", who);
    MB_CHECK_ARG(train_length >= 1 && first_step >= 0 && adam_steps_done >= 0, "%s: bad schedule", who);
    MB_CHECK_ARG(index == nullptr || index_stride >= train_length, "%s: index_stride < train_length", who);
    MB_CHECK_ARG(0 <= member_begin && member_begin <= member_end && member_end <= m->mlp.ensemble_size,
                 "%s: member slice [%d,%d) invalid", who, member_begin, member_end);
    MB_CHECK_ARG(obs && act && delta && reward && ensemble_loss && loss_terms, "%s: null tensor", who);
    MB_CHECK_ARG(!cfg->ignore_terminal_state || done, "%s: ignore_terminal_state needs done", who);
    if (nsteps == 0 || member_begin == member_end) return MB_OK;      // an empty member slice is a no-op
    if (!fused_train_supported(m->mlp, member_end - member_begin)) {
        set_error("%s: layer widths above 256 are not covered by the fused training kernel", who);
        return MB_EUNSUPPORTED;
    }
    FusedArgs a{};
    fused_pe_args(a, m, cfg, params, exp_avg, exp_avg_sq, adam_steps_done, member_begin, member_end);
    a.obs = obs; a.actn = act; a.delta = delta; a.reward = reward; a.done = done;
    a.index = index;
    a.index_stride = index_stride;
    a.per_member_rows = 1;
    a.nsteps = nsteps;
    a.first_step = first_step;
    a.train_length = train_length;
    a.batch_size = batch_size;
    a.ensemble_loss = ensemble_loss;
    a.loss_terms = loss_terms;
    return launch_train_fused(m->mlp, a, workspace, workspace_bytes, (cudaStream_t)stream);
}

static int train_common(const char* who, const mb_pe_model* m, const mb_train_cfg* cfg, float* params,
                        float* exp_avg, float* exp_avg_sq, int64_t step, int32_t* step_counter, const float* obs,
                        const float* act, const float* delta, const float* reward, const float* done,
                        int64_t B, int e0, int e1, float* grads, float* ensemble_loss,
                        float* loss_terms, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate_model(m, who);
    if (rc) return rc;
    MB_CHECK_ARG(cfg != nullptr, "%s: null cfg", who);
    MB_CHECK_ARG(m->mlp.activation == MB_ACT_SWISH, "%s: only swish models are trainable", who);
    MB_CHECK_ARG(B >= 1 && B < (int64_t)1 << 24, "%s: B %lld out of range", who, (long long)B);
    MB_CHECK_ARG(0 <= e0 && e0 <= e1 && e1 <= m->mlp.ensemble_size, "%s: member slice [%d,%d) invalid", who, e0,
                 e1);
    if (e0 == e1) return MB_OK;            // an empty slice (E = 7 members over 8 ranks) is a no-op
    MB_CHECK_ARG(obs && act && delta && reward && ensemble_loss && loss_terms, "%s: null tensor", who);
    MB_CHECK_ARG(!cfg->ignore_terminal_state || done, "%s: ignore_terminal_state needs done", who);
    if (workspace == nullptr || workspace_bytes < mb_pe_train_workspace_bytes(m, B)) {
        set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, mb_pe_train_workspace_bytes(m, B));
        return MB_EWORKSPACE;
    }
    if (params != nullptr && grads == nullptr && step_counter == nullptr && fused_enabled() &&
        fused_train_supported(m->mlp, e1 - e0)) {
        // ONE persistent launch: forward + NLL + backward + Adam (train_fused.cu)
        FusedArgs a{};
        fused_pe_args(a, m, cfg, params, exp_avg, exp_avg_sq, step - 1, e0, e1);
        a.obs = obs; a.actn = act; a.delta = delta; a.reward = reward; a.done = done;
        a.index = nullptr;
        a.index_stride = 0;
        a.per_member_rows = 1;
        a.nsteps = 1;
        a.first_step = 0;
        a.train_length = B;
        a.batch_size = B;
        a.ensemble_loss = ensemble_loss;
        a.loss_terms = loss_terms;
        return launch_train_fused(m->mlp, a, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    return launch_pe_train(to_dev(m), const_cast<void*>(m->tc_pack), *cfg, params, exp_avg, exp_avg_sq, step, step_counter, obs, act, delta, reward, done,
                           B, e0, e1, grads, ensemble_loss, loss_terms, workspace, (cudaStream_t)stream);
}

int mb_pe_train_step(const mb_pe_model* m, const mb_train_cfg* cfg, float* params, float* exp_avg,
                     float* exp_avg_sq, int64_t step, int32_t* step_counter, const float* obs, const float* act,
                     const float* delta, const float* reward, const float* done, int64_t B,
                     int member_begin, int member_end, float* ensemble_loss, float* loss_terms,
                     void* workspace, size_t workspace_bytes, void* stream) {
    MB_CHECK_ARG(params && exp_avg && exp_avg_sq, "mb_pe_train_step: null optimiser state");
    MB_CHECK_ARG(m && params == m->params, "mb_pe_train_step: params must be the model's blob");
    MB_CHECK_ARG(step >= 1 || step_counter, "mb_pe_train_step: step must be >= 1 (or pass step_counter)");
    return train_common("mb_pe_train_step", m, cfg, params, exp_avg, exp_avg_sq, step, step_counter, obs, act, delta,
                        reward, done, B, member_begin, member_end, nullptr, ensemble_loss, loss_terms,
                        workspace, workspace_bytes, stream);
}

int mb_pe_loss_grads(const mb_pe_model* m, const mb_train_cfg* cfg, const float* obs, const float* act,
                     const float* delta, const float* reward, const float* done, int64_t B,
                     float* grads, float* ensemble_loss, float* loss_terms, void* workspace,
                     size_t workspace_bytes, void* stream) {
    MB_CHECK_ARG(grads != nullptr, "mb_pe_loss_grads: null grads");
    return train_common("mb_pe_loss_grads", m, cfg, nullptr, nullptr, nullptr, 1, nullptr, obs, act, delta, reward,
                        done, B, 0, m ? m->mlp.ensemble_size : 0, grads, ensemble_loss, loss_terms, workspace,
                        workspace_bytes, stream);
}

/* ---- critic update: DDPGAgent.compute_qf_loss + qf_optimizer.step + soft target update ------------------ */
size_t mb_critic_train_workspace_bytes(const mb_mlp_desc* d, int64_t B) {
    if (!d || validate_desc(d, "mb_critic_train_workspace_bytes")) return 0;
    B = B < 1 ? 1 : B;
    size_t n = critic_chain_workspace_bytes(*d, B);
    if (fused_train_supported(*d, 1)) {
        const size_t f = fused_train_workspace_bytes(*d, B);
        n = f > n ? f : n;
    }
    return n;
}

int mb_critic_train_step(const mb_mlp_desc* d, float* params, float* exp_avg, float* exp_avg_sq,
                         int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                         const float* obs, const float* act, int obs_dim, int act_dim, const float* q_target,
                         int64_t B, float clip_error, float* target_params, float tau, float* ensemble_loss,
                         float* loss_terms, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "mb_critic_train_step";
    int rc = validate_desc(d, who);
    if (rc) return rc;
    MB_CHECK_ARG(obs_dim >= 1 && act_dim >= 0 && d->sizes[0] == obs_dim + act_dim, "%s: mlp input %d != obs+act %d", who,
                 d->sizes[0], obs_dim + act_dim);
    MB_CHECK_ARG(B >= 1 && B < (int64_t)1 << 20, "%s: B out of range", who);
    MB_CHECK_ARG(params && exp_avg && exp_avg_sq && obs && act && q_target && ensemble_loss && loss_terms,
                 "%s: null tensor", who);
    MB_CHECK_ARG(adam_steps_done >= 0, "%s: adam_steps_done < 0", who);
    if (!fused_train_supported(*d, d->ensemble_size)) {     // wide critics: per-layer launches
        MB_CHECK_ARG(workspace && workspace_bytes >= critic_chain_workspace_bytes(*d, B), "%s: workspace too small", who);
        return launch_critic_train_chain(*d, CRITIC_HEAD_MSE, params, exp_avg, exp_avg_sq, adam_steps_done, lr, beta1,
                                         beta2, eps, obs, act, obs_dim, act_dim, q_target, d->sizes[d->num_layers], B,
                                         clip_error, target_params, tau, ensemble_loss, loss_terms, workspace,
                                         (cudaStream_t)stream);
    }
    FusedArgs a{};
    a.e0 = 0;
    a.ne = d->ensemble_size;
    a.O = obs_dim;
    a.A = act_dim;
    a.D = 0;
    a.head_kind = FUSED_HEAD_MSE;
    a.E_total = d->ensemble_size;
    a.norm = nullptr;
    a.params = params;
    a.exp_avg = exp_avg;
    a.exp_avg_sq = exp_avg_sq;
    a.target_params = target_params;
    a.tau = tau;
    a.target = q_target;
    a.clip_error = clip_error;
    a.cfg = mb_train_cfg{};
    a.cfg.lr = lr; a.cfg.beta1 = beta1; a.cfg.beta2 = beta2; a.cfg.eps = eps;
    a.beta1_pow0 = pow((double)beta1, (double)adam_steps_done);
    a.beta2_pow0 = pow((double)beta2, (double)adam_steps_done);
    for (int l = 0; l < d->num_layers; ++l) a.ly[l].wd = 0.f;
    a.obs = obs; a.actn = act; a.delta = nullptr; a.reward = nullptr; a.done = nullptr;
    a.index = nullptr;
    a.index_stride = 0;
    a.per_member_rows = 0;                 // one batch shared by all members (ensemble_value.py:59-63,75-77)
    a.nsteps = 1;
    a.first_step = 0;
    a.train_length = B;
    a.batch_size = B;
    a.ensemble_loss = ensemble_loss;
    a.loss_terms = loss_terms;
    return launch_train_fused(*d, a, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mb_tqc_train_step(const mb_mlp_desc* d, float* params, float* exp_avg, float* exp_avg_sq,
                      int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                      const float* obs, const float* act, int obs_dim, int act_dim, const float* atoms_target,
                      int n_target, int64_t B, float* target_params, float tau, float* ensemble_loss,
                      float* loss_terms, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "mb_tqc_train_step";
    int rc = validate_desc(d, who);
    if (rc) return rc;
    MB_CHECK_ARG(obs_dim >= 1 && act_dim >= 0 && d->sizes[0] == obs_dim + act_dim, "%s: mlp input %d != obs+act %d", who,
                 d->sizes[0], obs_dim + act_dim);
    MB_CHECK_ARG(B >= 1 && B < (int64_t)1 << 20, "%s: B out of range", who);
    MB_CHECK_ARG(n_target >= 1, "%s: n_target < 1", who);
    MB_CHECK_ARG(params && exp_avg && exp_avg_sq && obs && act && atoms_target && ensemble_loss && loss_terms,
                 "%s: null tensor", who);
    MB_CHECK_ARG(adam_steps_done >= 0, "%s: adam_steps_done < 0", who);
    MB_CHECK_ARG(workspace && workspace_bytes >= critic_chain_workspace_bytes(*d, B), "%s: workspace too small", who);
    return launch_critic_train_chain(*d, CRITIC_HEAD_QUANTILE, params, exp_avg, exp_avg_sq, adam_steps_done, lr, beta1,
                                     beta2, eps, obs, act, obs_dim, act_dim, atoms_target, n_target, B, 0.f,
                                     target_params, tau, ensemble_loss, loss_terms, workspace, (cudaStream_t)stream);
}

size_t mb_policy_workspace_bytes(const mb_mlp_desc* d, int64_t B) {
    if (!d) return 0;
    if (B < 1) B = 1;
    const size_t fwd = critic_workspace_bytes(*d, B), pack = mlp_rows_pack_bytes(*d);
    return (((size_t)B * d->sizes[d->num_layers] * sizeof(float) + 255) & ~size_t(255)) + (fwd > pack ? fwd : pack) + 256;
}

int mb_policy_action(const mb_mlp_desc* d, const float* params, const float* obs, int obs_dim, int64_t B,
                     const float* eps, const mb_policy_cfg* cfg, float* action, float* log_prob, float* mean,
                     float* std, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "mb_policy_action";
    int rc = validate_desc(d, who);
    if (rc) return rc;
    MB_CHECK_ARG(d->ensemble_size == 1, "%s: the policy is a single network (ensemble_size %d)", who, d->ensemble_size);
    MB_CHECK_ARG(obs_dim >= 1 && d->sizes[0] == obs_dim, "%s: mlp input %d != obs_dim %d", who, d->sizes[0], obs_dim);
    MB_CHECK_ARG(d->sizes[d->num_layers] % 2 == 0, "%s: output width %d is not 2 * act_dim", who, d->sizes[d->num_layers]);
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "%s: B out of range", who);
    MB_CHECK_ARG(cfg != nullptr && cfg->bound_mode >= 0 && cfg->bound_mode <= 2, "%s: bad cfg", who);
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(params && obs && action, "%s: null tensor", who);
    const size_t need = ((size_t)B * d->sizes[d->num_layers] * sizeof(float) + 255) & ~size_t(255);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, need);
        return MB_EWORKSPACE;
    }
    return launch_policy_action(*d, params, obs, obs_dim, B, eps, *cfg, action, log_prob, mean, std, workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

size_t mb_actor_workspace_bytes(const mb_mlp_desc* policy, const mb_mlp_desc* critic, int64_t B) {
    if (!policy || !critic || validate_desc(policy, "mb_actor_workspace_bytes") || validate_desc(critic, "mb_actor_workspace_bytes"))
        return 0;
    return actor_workspace_bytes(*policy, *critic, B < 1 ? 1 : B);
}

int mb_actor_update(const mb_mlp_desc* policy, float* policy_params, float* exp_avg, float* exp_avg_sq,
                    int64_t adam_steps_done, float lr, float beta1, float beta2, float eps,
                    const mb_policy_cfg* cfg, const mb_mlp_desc* critic, const float* critic_params,
                    int reduce_mode, const int32_t* members, int n_members,
                    const float* obs, const float* noise, int obs_dim, int act_dim, int64_t B,
                    float* log_alpha, float* alpha_exp_avg, float* alpha_exp_avg_sq, float target_entropy,
                    float alpha_fixed, float* info, void* workspace, size_t workspace_bytes, void* stream) {
    const char* who = "mb_actor_update";
    int rc = validate_desc(policy, who);
    if (rc) return rc;
    rc = validate_desc(critic, who);
    if (rc) return rc;
    MB_CHECK_ARG(cfg != nullptr, "%s: null policy cfg", who);
    MB_CHECK_ARG(policy->ensemble_size == 1, "%s: the policy is a 1-member MLP", who);
    MB_CHECK_ARG(obs_dim >= 1 && act_dim >= 1 && policy->sizes[0] == obs_dim &&
                 policy->sizes[policy->num_layers] == 2 * act_dim, "%s: policy maps %d -> %d, expected %d -> %d", who,
                 policy->sizes[0], policy->sizes[policy->num_layers], obs_dim, 2 * act_dim);
    MB_CHECK_ARG(critic->sizes[0] == obs_dim + act_dim, "%s: critic input %d != obs+act %d", who, critic->sizes[0],
                 obs_dim + act_dim);
    MB_CHECK_ARG(B >= 1 && B < (int64_t)1 << 20, "%s: B out of range", who);
    MB_CHECK_ARG(policy_params && exp_avg && exp_avg_sq && critic_params && obs && noise, "%s: null tensor", who);
    MB_CHECK_ARG(log_alpha == nullptr || (alpha_exp_avg && alpha_exp_avg_sq), "%s: log_alpha without its Adam moments", who);
    MB_CHECK_ARG(adam_steps_done >= 0, "%s: adam_steps_done < 0", who);
    MB_CHECK_ARG(reduce_mode >= MB_REDUCE_MIN && reduce_mode <= MB_REDUCE_MAX, "%s: reduce_mode %d", who, reduce_mode);
    const int Q = critic->sizes[critic->num_layers];
    MB_CHECK_ARG(Q == 1 || reduce_mode == MB_REDUCE_MEAN, "%s: critics with %d outputs reduce by the mean of their atoms", who, Q);
    MB_CHECK_ARG(members == nullptr || (n_members >= 1 && n_members <= critic->ensemble_size), "%s: n_members %d", who, n_members);
    for (int i = 0; members && i < n_members; ++i)
        MB_CHECK_ARG(members[i] >= 0 && members[i] < critic->ensemble_size, "%s: member %d out of range", who, members[i]);
    MB_CHECK_ARG(workspace && workspace_bytes >= actor_workspace_bytes(*policy, *critic, B), "%s: workspace too small", who);
    return launch_actor_update(*policy, policy_params, exp_avg, exp_avg_sq, adam_steps_done, lr, beta1, beta2, eps, *cfg,
                               *critic, critic_params, reduce_mode, members, n_members, obs, noise, obs_dim, act_dim, B,
                               log_alpha, alpha_exp_avg, alpha_exp_avg_sq, target_entropy, alpha_fixed, info, workspace,
                               (cudaStream_t)stream);
}

static int check_critic(const char* who, const mb_mlp_desc* d, const float* params, const float* obs,
                        const float* act, int O, int A, int64_t B) {
    int rc = validate_desc(d, who);
    if (rc) return rc;
    MB_CHECK_ARG(O >= 1 && A >= 0 && d->sizes[0] == O + A, "%s: mlp input %d != obs+act %d", who, d->sizes[0],
                 O + A);
    MB_CHECK_ARG(B >= 0 && B < (int64_t)1 << 31, "%s: B out of range", who);
    MB_CHECK_ARG(params && obs && act, "%s: null tensor", who);
    return MB_OK;
}

size_t mb_critic_workspace_bytes(const mb_mlp_desc* d, int64_t B) {
    return d ? critic_workspace_bytes(*d, B < 1 ? 1 : B) : 0;
}

// a workspace that is too small is treated like no workspace (single-launch tile kernels)
static void* usable_ws(const mb_mlp_desc* d, int64_t B, void* ws, size_t bytes) {
    return (ws && bytes >= critic_workspace_bytes(*d, B)) ? ws : nullptr;
}

int mb_critic_forward(const mb_mlp_desc* d, const float* params, const float* obs, const float* act,
                      int obs_dim, int act_dim, int64_t B, float* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
    int rc = check_critic("mb_critic_forward", d, params, obs, act, obs_dim, act_dim, B);
    if (rc) return rc;
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(out != nullptr, "mb_critic_forward: null out");
    return launch_critic_forward(*d, params, obs, act, obs_dim, act_dim, B, out,
                                 usable_ws(d, B, workspace, workspace_bytes), (cudaStream_t)stream);
}

int mb_critic_redq_target(const mb_mlp_desc* d, const float* params, const float* obs, const float* act,
                          int obs_dim, int act_dim, int64_t B, const int32_t* members_host,
                          const uint8_t* row_members, int n_sub, int mode, int td, const float* reward,
                          const float* terminal, const float* log_prob, float alpha, float discount,
                          float* value, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_critic("mb_critic_redq_target", d, params, obs, act, obs_dim, act_dim, B);
    if (rc) return rc;
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(value != nullptr, "mb_critic_redq_target: null value");
    MB_CHECK_ARG(d->sizes[d->num_layers] == 1, "mb_critic_redq_target: critic output must be 1");
    MB_CHECK_ARG(mode >= MB_REDUCE_MIN && mode <= MB_REDUCE_MAX, "mb_critic_redq_target: unknown mode %d", mode);
    MB_CHECK_ARG(!(members_host && row_members), "mb_critic_redq_target: pass one of members_host/row_members");
    if (members_host || row_members)
        MB_CHECK_ARG(n_sub >= 1 && n_sub <= d->ensemble_size, "mb_critic_redq_target: n_sub %d out of range", n_sub);
    if (members_host)
        for (int i = 0; i < n_sub; ++i)
            MB_CHECK_ARG(members_host[i] >= 0 && members_host[i] < d->ensemble_size,
                         "mb_critic_redq_target: member id %d out of range", members_host[i]);
    MB_CHECK_ARG(!td || (reward && terminal && log_prob), "mb_critic_redq_target: td needs reward/terminal/log_prob");
    return launch_critic_redq(*d, params, obs, act, obs_dim, act_dim, B, members_host, row_members, n_sub, mode,
                              td, reward, terminal, log_prob, alpha, discount, value,
                              usable_ws(d, B, workspace, workspace_bytes), (cudaStream_t)stream);
}

int mb_critic_tqc_target(const mb_mlp_desc* d, const float* params, const float* obs, const float* act,
                         int obs_dim, int act_dim, int64_t B, int n_drop, int td, const float* reward,
                         const float* terminal, const float* log_prob, float alpha, float discount,
                         float* value, float* atoms, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_critic("mb_critic_tqc_target", d, params, obs, act, obs_dim, act_dim, B);
    if (rc) return rc;
    if (B == 0) return MB_OK;
    const int NA = d->ensemble_size * d->sizes[d->num_layers];
    MB_CHECK_ARG(n_drop >= 0 && n_drop < NA, "mb_critic_tqc_target: n_drop %d outside [0,%d)", n_drop, NA);
    MB_CHECK_ARG(NA <= 512, "mb_critic_tqc_target: %d atoms per row unsupported (max 512)", NA);
    MB_CHECK_ARG(!td || (reward && terminal && log_prob), "mb_critic_tqc_target: td needs reward/terminal/log_prob");
    return launch_critic_tqc(*d, params, obs, act, obs_dim, act_dim, B, n_drop, td, reward, terminal, log_prob,
                             alpha, discount, value, atoms, usable_ws(d, B, workspace, workspace_bytes),
                             (cudaStream_t)stream);
}

size_t mb_pe_draw_members_workspace_bytes(int64_t B, int n_elites) {
    if (B <= 0 || n_elites < 1 || n_elites > 64) return 0;
    return draw_workspace_bytes(B, n_elites);
}

int mb_pe_draw_members_chained(uint32_t* mt_state, uint32_t* state_snapshot, const uint8_t* elites_host,
                               int n_elites, int64_t B, uint8_t* member_idx, void* workspace,
                               size_t workspace_bytes, void* stream) {
    MB_CHECK_ARG(mt_state && elites_host && n_elites >= 1 && n_elites <= 64,
                 "mb_pe_draw_members: need 1..64 elites and a state buffer (got %d)", n_elites);
    MB_CHECK_ARG(B >= 0 && B <= (int64_t)1 << 28, "mb_pe_draw_members: B %lld out of range", (long long)B);
    if (B == 0) return MB_OK;
    MB_CHECK_ARG(member_idx != nullptr, "mb_pe_draw_members: null output");
    if (workspace == nullptr || workspace_bytes < draw_workspace_bytes(B, n_elites)) {
        set_error("mb_pe_draw_members: workspace %zu < %zu bytes", workspace_bytes, draw_workspace_bytes(B, n_elites));
        return MB_EWORKSPACE;
    }
    return launch_draw_members(mt_state, state_snapshot, elites_host, n_elites, B, member_idx, workspace,
                               (cudaStream_t)stream);
}

int mb_pe_draw_members(uint32_t* mt_state, const uint8_t* elites_host, int n_elites, int64_t B,
                       uint8_t* member_idx, void* workspace, size_t workspace_bytes, void* stream) {
    return mb_pe_draw_members_chained(mt_state, nullptr, elites_host, n_elites, B, member_idx, workspace,
                                      workspace_bytes, stream);
}

// ---- peer-mapped pools (CUDA IPC): destinations of mb_pe_rollout_step_rows on other GPUs ---------
int mb_peer_pool_alloc(size_t bytes, void** dev_ptr, unsigned char* handle) {
    MB_CHECK_ARG(dev_ptr && handle && bytes > 0, "mb_peer_pool_alloc: null argument or zero size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) == MB_IPC_HANDLE_BYTES, "IPC handle size");
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(handle, &h, sizeof(h));
        else cudaFree(p);
    }
    if (e != cudaSuccess) {
        set_error("mb_peer_pool_alloc(%zu): %s", bytes, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MB_ECUDA;
    }
    *dev_ptr = p;
    return MB_OK;
}

int mb_peer_pool_open(const unsigned char* handle, void** dev_ptr) {
    MB_CHECK_ARG(dev_ptr && handle, "mb_peer_pool_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    // lazy peer access: the driver maps the exporting GPU's memory over NVLink / PCIe P2P
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("mb_peer_pool_open: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MB_ECUDA;
    }
    *dev_ptr = p;
    return MB_OK;
}

int mb_peer_pool_close(void* dev_ptr) {
    if (dev_ptr == nullptr) return MB_OK;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    if (e != cudaSuccess) {
        set_error("mb_peer_pool_close: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MB_ECUDA;
    }
    return MB_OK;
}

int mb_peer_pool_free(void* dev_ptr) {
    if (dev_ptr == nullptr) return MB_OK;
    cudaError_t e = cudaFree(dev_ptr);
    if (e != cudaSuccess) {
        set_error("mb_peer_pool_free: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MB_ECUDA;
    }
    return MB_OK;
}

int mb_set_reserved_sms(int n) {
    MB_CHECK_ARG(n >= 0 && n <= 64, "mb_set_reserved_sms: %d outside [0,64]", n);
    g_reserved_sms = n;
    return MB_OK;
}

int mb_profile_enable(int on) {
    if (on && g_prof_ev[0] == nullptr) {
        if (cudaEventCreate(&g_prof_ev[0]) != cudaSuccess || cudaEventCreate(&g_prof_ev[1]) != cudaSuccess) {
            set_error("mb_profile_enable: cudaEventCreate failed");
            return MB_ECUDA;
        }
    }
    g_prof_on = on != 0;
    return MB_OK;
}

int mb_profile_last_ms(float* ms) {
    MB_CHECK_ARG(ms != nullptr && g_prof_ev[0] != nullptr, "mb_profile_last_ms: profiling was never enabled");
    cudaError_t e = cudaEventSynchronize(g_prof_ev[1]);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, g_prof_ev[0], g_prof_ev[1]);
    if (e != cudaSuccess) { set_error("mb_profile_last_ms: %s", cudaGetErrorString(e)); return MB_ECUDA; }
    return MB_OK;
}

size_t mb_pe_tc_pack_bytes(const mb_pe_model* m) { return m ? tc_pack_bytes(m->mlp) : 0; }

int mb_pe_tc_pack(const mb_pe_model* m, void* pack, void* stream) {
    int rc = validate_model(m, "mb_pe_tc_pack");
    if (rc) return rc;
    MB_CHECK_ARG(pack != nullptr, "mb_pe_tc_pack: null pack");
    return launch_tc_pack(to_dev(m), pack, (cudaStream_t)stream);
}

}  // extern "C"
