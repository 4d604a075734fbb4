"""ctypes binding of ``libmirl_b200.so`` (C ABI declared in ``include/mirl_b200.h``).

There is no fallback: if the library has not been built, or a call fails, this raises.  Nothing
here (or anywhere in ``mirl_b200``) imports ``oracle``.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)

import torch

MB_MAX_LAYERS = 8
ACT = {"swish": 0, "relu": 1}
DONE_KIND = {"cheetah": 0, "hopper": 1, "walker2d": 2, "mb_ant": 3, "mb_humanoid": 4,
             "inverted_pendulum": 5}
PENALTY = {None: 0, "total_var": 1, "max_var": 2}
REDUCE = {"min": 0, "mean": 1, "max": 2}
KERNEL = {"simt": 0, "tcgen05": 1}

_HERE = os.path.dirname(os.path.abspath(__file__))
# MIRL_B200_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("MIRL_B200_LIB") or os.path.join(_HERE, "libmirl_b200.so")


class MlpDesc(Structure):
    _fields_ = [("ensemble_size", c_int32), ("num_layers", c_int32),
                ("sizes", c_int32 * (MB_MAX_LAYERS + 1)), ("activation", c_int32)]


class PeModel(Structure):
    _fields_ = [("mlp", MlpDesc), ("obs_dim", c_int32), ("act_dim", c_int32),
                ("params", c_void_p), ("norm", c_void_p), ("tc_pack", c_void_p)]


class TrainCfg(Structure):
    _fields_ = [("weight_decay", c_float * MB_MAX_LAYERS), ("bound_reg_coef", c_float),
                ("reward_coef", c_float), ("ignore_terminal_state", c_int32), ("lr", c_float),
                ("beta1", c_float), ("beta2", c_float), ("eps", c_float)]


class PolicyCfg(Structure):
    _fields_ = [("mean_scale", c_float), ("std_scale", c_float), ("min_std", c_float),
                ("logstd_extra_bias", c_float), ("bound_mode", c_int32), ("squashed", c_int32)]


LOGSTD_BOUND = {"tanh": 0, "clamp": 1, "no": 2}

# name -> (restype, argtypes); one entry per symbol include/mirl_b200.h declares
P = c_void_p
SIGNATURES = {
    "mb_abi_version": (c_int, []),
    "mb_last_error": (c_char_p, []),
    "mb_mlp_weight_offset": (c_int64, [POINTER(MlpDesc), c_int]),
    "mb_mlp_bias_offset": (c_int64, [POINTER(MlpDesc), c_int]),
    "mb_mlp_logstd_offset": (c_int64, [POINTER(MlpDesc), c_int]),
    "mb_mlp_param_count": (c_int64, [POINTER(MlpDesc), c_int]),
    "mb_pe_rollout_workspace_bytes": (c_size_t, [POINTER(PeModel), c_int64]),
    "mb_pe_rollout_step": (c_int, [POINTER(PeModel), c_int, P, P, P, P, c_int64, c_int, P, P, P,
                                   P, c_size_t, P]),
    "mb_pe_draw_members_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "mb_pe_draw_members": (c_int, [P, ctypes.c_char_p, c_int, c_int64, P, P, c_size_t, P]),
    "mb_pe_draw_members_chained": (c_int, [P, P, ctypes.c_char_p, c_int, c_int64, P, P, c_size_t, P]),
    "mb_pe_rollout_step_rows": (c_int, [POINTER(PeModel), c_int, P, c_int64, P, c_int64, P, P, c_int64, c_int,
                                        POINTER(c_void_p), c_int, c_int64, c_int64, c_int, P, c_size_t, P]),
    "mb_pe_dist_workspace_bytes": (c_size_t, [POINTER(PeModel), c_int64, c_int]),
    "mb_pe_rollout_step_dist_tc": (c_int, [POINTER(PeModel), P, P, P, P, c_int64, c_int, POINTER(c_int32), c_int,
                                           P, P, P, P, P, P, P, P, c_int, c_float, P, P, P, c_size_t, P]),
    "mb_pe_forward_tc": (c_int, [POINTER(PeModel), P, P, c_int, c_int64, POINTER(c_int32), c_int, P, P]),
    "mb_pe_rollout_step_dist": (c_int, [POINTER(PeModel), P, P, P, P, c_int64, c_int,
                                        POINTER(c_int32), c_int, P, P, P, P, P, P, P, P, c_int,
                                        c_float, P, P, P]),
    "mb_pe_forward": (c_int, [POINTER(PeModel), P, P, c_int, c_int, c_int64, P, P, P]),
    "mb_pe_loss": (c_int, [POINTER(PeModel), POINTER(TrainCfg), P, P, P, P, P, c_int, c_int64, P,
                           P, P]),
    "mb_pe_train_workspace_bytes": (c_size_t, [POINTER(PeModel), c_int64]),
    "mb_pe_train_step": (c_int, [POINTER(PeModel), POINTER(TrainCfg), P, P, P, c_int64, P, P, P, P,
                                 P, P, c_int64, c_int, c_int, P, P, P, c_size_t, P]),
    "mb_pe_train_steps_workspace_bytes": (c_size_t, [POINTER(PeModel), c_int64]),
    "mb_pe_train_steps": (c_int, [POINTER(PeModel), POINTER(TrainCfg), P, P, P, c_int64, P, P, P, P, P, P,
                                  c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "mb_pe_loss_grads": (c_int, [POINTER(PeModel), POINTER(TrainCfg), P, P, P, P, P, c_int64, P,
                                 P, P, P, c_size_t, P]),
    "mb_critic_train_workspace_bytes": (c_size_t, [POINTER(MlpDesc), c_int64]),
    "mb_critic_train_step": (c_int, [POINTER(MlpDesc), P, P, P, c_int64, c_float, c_float, c_float, c_float, P, P,
                                     c_int, c_int, P, c_int64, c_float, P, c_float, P, P, P, c_size_t, P]),
    "mb_tqc_train_step": (c_int, [POINTER(MlpDesc), P, P, P, c_int64, c_float, c_float, c_float, c_float, P, P,
                                  c_int, c_int, P, c_int, c_int64, P, c_float, P, P, P, c_size_t, P]),
    "mb_policy_workspace_bytes": (c_size_t, [POINTER(MlpDesc), c_int64]),
    "mb_policy_action": (c_int, [POINTER(MlpDesc), P, P, c_int, c_int64, P, POINTER(PolicyCfg), P, P, P, P, P, c_size_t, P]),
    "mb_actor_workspace_bytes": (c_size_t, [POINTER(MlpDesc), POINTER(MlpDesc), c_int64]),
    "mb_actor_update": (c_int, [POINTER(MlpDesc), P, P, P, c_int64, c_float, c_float, c_float, c_float, POINTER(PolicyCfg),
                                POINTER(MlpDesc), P, c_int, POINTER(c_int32), c_int, P, P, c_int, c_int, c_int64,
                                P, P, P, c_float, c_float, P, P, c_size_t, P]),
    "mb_critic_workspace_bytes": (c_size_t, [POINTER(MlpDesc), c_int64]),
    "mb_critic_forward": (c_int, [POINTER(MlpDesc), P, P, P, c_int, c_int, c_int64, P, P, c_size_t, P]),
    "mb_critic_redq_target": (c_int, [POINTER(MlpDesc), P, P, P, c_int, c_int, c_int64,
                                      POINTER(c_int32), P, c_int, c_int, c_int, P, P, P, c_float,
                                      c_float, P, P, c_size_t, P]),
    "mb_critic_tqc_target": (c_int, [POINTER(MlpDesc), P, P, P, c_int, c_int, c_int64, c_int,
                                     c_int, P, P, P, c_float, c_float, P, P, P, c_size_t, P]),
    "mb_pe_tc_pack_bytes": (c_size_t, [POINTER(PeModel)]),
    "mb_pe_tc_pack": (c_int, [POINTER(PeModel), P, P]),
    "mb_peer_pool_alloc": (c_int, [c_size_t, POINTER(c_void_p), ctypes.c_char_p]),
    "mb_peer_pool_open": (c_int, [ctypes.c_char_p, POINTER(c_void_p)]),
    "mb_peer_pool_close": (c_int, [P]),
    "mb_peer_pool_free": (c_int, [P]),
    "mb_set_reserved_sms": (c_int, [c_int]),
    "mb_profile_enable": (c_int, [c_int]),
    "mb_profile_last_ms": (c_int, [POINTER(c_float)]),
}

_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "mirl_b200: %s is missing -- build it with `python -m mirl_b200.build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)            # AttributeError if a symbol is missing
            fn.restype, fn.argtypes = res, args
        if handle.mb_abi_version() != 1:
            raise RuntimeError("mirl_b200: ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("mirl_b200 error %d: %s" % (rc, lib().mb_last_error().decode()))


def make_desc(sizes, ensemble_size, activation):
    if activation not in ACT:
        raise NotImplementedError("activation %r (swish and relu are the ones the path uses)"
                                  % (activation,))
    if len(sizes) - 1 > MB_MAX_LAYERS:
        raise NotImplementedError("more than %d layers" % MB_MAX_LAYERS)
    d = MlpDesc()
    d.ensemble_size = int(ensemble_size)
    d.num_layers = len(sizes) - 1
    for i, s in enumerate(sizes):
        d.sizes[i] = int(s)
    d.activation = ACT[activation]
    return d


def ptr(t):
    """Device pointer of a CUDA fp32/uint8 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mirl_b200 kernels need CUDA tensors (got a %s tensor); there is no "
                           "CPU path" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("mirl_b200: tensor must be contiguous")
    return c_void_p(t.data_ptr())


def stream(device=None):
    """The CUDA stream the launches go to: torch's current stream of ``device`` (the tensors'
    device, not whatever device happens to be current)."""
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(fn):
    """Run a method with its object's CUDA device current: the C ABI launches on the CURRENT device's
    stream and caches per-device kernel attributes, so a model on cuda:1 must not launch while cuda:0 is
    current (round-1 advisor finding)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        dev = self.device
        if getattr(dev, "type", None) != "cuda":
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper


def f32(t, device):
    """Contiguous fp32 tensor on ``device`` (no copy when it already is one)."""
    return t.to(device=device, dtype=torch.float32).contiguous()
