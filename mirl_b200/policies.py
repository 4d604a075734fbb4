"""B200GaussianPolicy -- drop-in for ``GaussianPolicy`` (mirl/policies/gaussian_policy.py:8-67) over
``MeanLogstdGaussian`` (mirl/torch_modules/gaussian.py:170-204) for the rollout loop.

``MBPOCollector.imagine`` calls ``policy.action(o)`` once per horizon step (mbpo_collector.py:48): in the
reference that is an eager 17 -> 256 -> 256 -> 12 MLP, ``chunk``, ``tanh`` bound, ``exp``, ``Normal.rsample``,
``tanh`` -- about 15 ATen kernels between two launches of the fused rollout step.  Here it is the layer GEMMs
of the ensemble kernels (a 1-member "ensemble") plus ONE head kernel (``mb_policy_action``): log-std bound,
reparameterised sample from a caller-visible noise tensor, tanh squashing and the log-probability.

Same constructor arguments, ``action()`` keywords, return values and state-dict keys
(``module.net.{2l}.weight`` ``[in,out]``, ``module.net.{2l}.bias`` ``[1,out]``) as the reference, so reference
checkpoints load unchanged.  Forward-only (the SAC actor update differentiates through the reference policy;
this class serves the no-grad uses: model rollouts, exploration, target actions).  The noise is drawn like the
reference's ``dist.rsample()`` / ``dist.sample()``: one standard-normal ``[B, A]`` tensor from the default
generator of the observation's device; ``noise=`` replays a recorded draw.
"""
import ctypes
from contextlib import contextmanager

import torch
import torch.nn as nn

from . import _lib
from .compat import Policy, snapshot_dir
from .ensemble import PackedEnsembleMLP


class B200GaussianPolicy(nn.Module, Policy):
    def __init__(self, env, deterministic=False, squashed=True, policy_name="gaussian_policy", bound_mode="tanh",
                 logstd_extra_bias=0, mean_scale=1.0, std_scale=1.0, min_std=1e-4, **mlp_kwargs):
        nn.Module.__init__(self)
        Policy.__init__(self, env)
        if bound_mode not in _lib.LOGSTD_BOUND:
            raise NotImplementedError("bound_mode %r" % (bound_mode,))
        self._deterministic = deterministic
        self.squashed = squashed
        self.bound_mode, self.logstd_extra_bias = bound_mode, logstd_extra_bias
        self.mean_scale, self.std_scale, self.min_std = mean_scale, std_scale, min_std
        O, A = self.observation_shape[0], self.action_shape[0]
        mlp_kwargs.pop("init_func_name", None)
        self.net = PackedEnsembleMLP(O, 2 * A, mlp_kwargs.pop("hidden_layers"), 1, module_name=policy_name,
                                     **mlp_kwargs)
        self._ws = None

    @contextmanager
    def deterministic_(self, deterministic=True):
        """base_policy.py:42-54."""
        was = self._deterministic
        self._deterministic = deterministic
        yield
        self._deterministic = was

    @property
    def device(self):
        return self.net.flat.device

    # ---- reference state-dict keys ------------------------------------------------------------------
    def _reference_items(self):
        for l in range(self.net.num_fc):
            yield "module.net.%d.weight" % (2 * l), self.net.weight(l)[0]
            yield "module.net.%d.bias" % (2 * l), self.net.bias(l)[0].view(1, -1)

    def state_dict(self, *args, **kwargs):
        return {k: v.detach().clone() for k, v in self._reference_items()}

    def load_state_dict(self, sd, strict=True):
        own = dict(self._reference_items())
        missing = [k for k in own if k not in sd]
        if strict and (missing or [k for k in sd if k not in own]):
            raise RuntimeError("state dict mismatch: missing %s" % missing)
        with torch.no_grad():
            for k, view in own.items():
                if k in sd:
                    view.copy_(torch.as_tensor(sd[k]).reshape(view.shape))
        self.net.mark_weights_changed()

    def get_snapshot(self):
        return self.state_dict()

    def save(self, save_dir=None):
        import os.path as osp
        torch.save(self.state_dict(), osp.join(save_dir or snapshot_dir(), "%s.pt" % self.net.module_name))

    def load(self, load_dir=None):
        import os.path as osp
        self.load_state_dict(torch.load(osp.join(load_dir or snapshot_dir(), "%s.pt" % self.net.module_name)))

    # ---- action ---------------------------------------------------------------------------------------
    def action(self, obs, return_log_prob=False, reparameterize=True, return_mean_std=False, return_dist=False,
               noise=None, **kwargs):
        """``GaussianPolicy.action`` (gaussian_policy.py:35-47) -> ``Gaussian.forward`` (gaussian.py:55-97):
        ``(action [B,A], info)`` with ``info['log_prob'] [B,1]`` / ``info['mean'], info['std']`` on request."""
        if return_dist or kwargs:
            raise NotImplementedError("B200GaussianPolicy.action options %s" % (sorted(kwargs) + ["return_dist"] * return_dist))
        dev = self.device
        L = _lib.lib()
        obs = _lib.f32(obs, dev)
        assert obs.dim() == 2
        B, A = obs.shape[0], self.action_shape[0]
        if self._deterministic:
            noise = None
        elif noise is None:
            noise = torch.randn(B, A, device=dev)          # Normal.rsample()/sample(): one [B,A] standard-normal draw
        else:
            noise = _lib.f32(noise, dev)
            assert noise.shape == (B, A)
        action = torch.empty(B, A, device=dev)
        info = {}
        lp = torch.empty(B, 1, device=dev) if return_log_prob else None
        mean = torch.empty(B, A, device=dev) if return_mean_std else None
        std = torch.empty(B, A, device=dev) if return_mean_std else None
        if B > 0:
            need = L.mb_policy_workspace_bytes(ctypes.byref(self.net.desc), B)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=dev)
            cfg = _lib.PolicyCfg(float(self.mean_scale), float(self.std_scale), float(self.min_std),
                                 float(self.logstd_extra_bias), _lib.LOGSTD_BOUND[self.bound_mode], int(bool(self.squashed)))
            with torch.cuda.device(dev):
                _lib.check(L.mb_policy_action(
                    ctypes.byref(self.net.desc), _lib.ptr(self.net.flat.data), _lib.ptr(obs), self.observation_shape[0], B,
                    _lib.ptr(noise), ctypes.byref(cfg), _lib.ptr(action), _lib.ptr(lp), _lib.ptr(mean), _lib.ptr(std),
                    _lib.ptr(self._ws), self._ws.numel(), _lib.stream(dev)))
        if return_log_prob:
            info["log_prob"] = lp
        if return_mean_std:
            info["mean"], info["std"] = mean, std
        return action, info

    def action_np(self, obs, **kwargs):
        a, info = self.action(torch.as_tensor(obs, device=self.device).float(), **kwargs)
        return a.cpu().numpy(), {k: v.cpu().numpy() for k, v in info.items()}
