// Shared host/device helpers for the mirl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/mirl_b200.h"

namespace mb {

void set_error(const char* fmt, ...);

#define MB_CHECK_ARG(cond, ...)                          \
    do {                                                 \
        if (!(cond)) {                                   \
            mb::set_error(__VA_ARGS__);                  \
            return MB_EINVAL;                            \
        }                                                \
    } while (0)

#define MB_CUDA_LAUNCH_CHECK(what)                                                   \
    do {                                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) {                                                     \
            mb::set_error("%s: %s", what, cudaGetErrorString(_e));                   \
            return MB_ECUDA;                                                         \
        }                                                                            \
    } while (0)

__host__ __device__ inline int64_t round_up4(int64_t x) { return (x + 3) & ~int64_t(3); }

// ---- programmatic dependent launch (the latency-bound chains: train step, critic layers) -------
// A kernel launched through launch_pdl may become resident while its predecessor in the stream is
// still running (as soon as every CTA of the predecessor has called pdl_trigger or exited); it
// runs its prologue (barrier init, TMEM allocation, descriptor copies) and then blocks in
// pdl_wait until the predecessor has completed and its writes are visible.  Both instructions
// are no-ops for a grid launched the ordinary way.  MB_PDL=0 launches everything the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);      // errors surface in MB_CUDA_LAUNCH_CHECK
}
#endif

// ---- packed parameter blob layout (include/mirl_b200.h, mb_mlp_desc) ----------------------
__host__ __device__ inline int64_t w_offset(const mb_mlp_desc& d, int layer) {
    int64_t off = 0;
    for (int l = 0; l < layer; ++l) {
        off = round_up4(off + (int64_t)d.ensemble_size * d.sizes[l] * d.sizes[l + 1]);
        off = round_up4(off + (int64_t)d.ensemble_size * d.sizes[l + 1]);
    }
    return off;
}
__host__ __device__ inline int64_t b_offset(const mb_mlp_desc& d, int layer) {
    return round_up4(w_offset(d, layer) +
                     (int64_t)d.ensemble_size * d.sizes[layer] * d.sizes[layer + 1]);
}
__host__ __device__ inline int64_t logstd_offset(const mb_mlp_desc& d, int which) {
    int64_t off = w_offset(d, d.num_layers);
    if (which) off = round_up4(off + (int64_t)d.ensemble_size * (d.sizes[d.num_layers] / 2));
    return off;
}
__host__ __device__ inline int64_t param_count(const mb_mlp_desc& d, bool probabilistic) {
    if (!probabilistic) return w_offset(d, d.num_layers);
    return round_up4(logstd_offset(d, 1) + (int64_t)d.ensemble_size * (d.sizes[d.num_layers] / 2));
}

int validate_desc(const mb_mlp_desc* d, const char* who);
int validate_model(const mb_pe_model* m, const char* who);

// ---- scalar maths, written to follow the ATen fp32 kernels the reference dispatches -------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float swishf_(float x) { return x * sigmoidf_(x); }
// F.softplus, beta 1, threshold 20 (pe_model.py:149-150)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
// d softplus / dx with the same threshold (ATen softplus_backward)
__device__ __forceinline__ float softplus_gradf_(float x) {
    return x > 20.0f ? 1.0f : sigmoidf_(x);
}
// PEModelModule._get_mean_logstd, pe_model.py:145-151
__device__ __forceinline__ float clamp_logstd(float raw, float mx, float mn) {
    float ls = mx - softplusf_(mx - raw);
    return mn + softplusf_(ls - mn);
}
__device__ __forceinline__ float apply_act(float x, int act) {
    return act == MB_ACT_RELU ? fmaxf(x, 0.0f) : swishf_(x);
}

// done functions, mirl/environments/reward_done_functions/*.py (one row, O observations)
__device__ __forceinline__ float done_fn(int kind, const float* next_obs, int O) {
    switch (kind) {
    case MB_DONE_CHEETAH: return 0.0f;
    case MB_DONE_HOPPER: {
        bool ok = true;
        for (int j = 0; j < O; ++j) ok = ok && isfinite(next_obs[j]);
        for (int j = 1; j < O; ++j) ok = ok && (fabsf(next_obs[j]) < 100.0f);
        ok = ok && (next_obs[0] > 0.7f) && (fabsf(next_obs[1]) < 0.2f);
        return ok ? 0.0f : 1.0f;
    }
    case MB_DONE_WALKER2D: {
        float h = next_obs[0], a = next_obs[1];
        bool live = (h > 0.8f) && (h < 2.0f) && (a > -1.0f) && (a < 1.0f);
        return live ? 0.0f : 1.0f;
    }
    case MB_DONE_MB_ANT: {
        bool ok = true;
        for (int j = 0; j < O; ++j) ok = ok && isfinite(next_obs[j]);
        ok = ok && (next_obs[0] >= 0.2f) && (next_obs[0] <= 1.0f);
        return ok ? 0.0f : 1.0f;
    }
    case MB_DONE_MB_HUMANOID: {
        float q = next_obs[0];
        return ((q < 1.0f) || (q > 2.0f)) ? 1.0f : 0.0f;
    }
    case MB_DONE_INVERTED_PENDULUM: {
        bool ok = true;
        for (int j = 0; j < O; ++j) ok = ok && isfinite(next_obs[j]);
        ok = ok && (next_obs[1] <= 0.2f) && (next_obs[1] >= -0.2f);
        return ok ? 0.0f : 1.0f;
    }
    }
    return 0.0f;
}

constexpr float kNormEps = 1e-6f;              // Normalizer epsilon, normalizer.py:7
constexpr float kLog2Pi = 1.8378770664093453f; // pe_model.py:13

// ---- cp.async (LDGSTS) helpers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace mb
