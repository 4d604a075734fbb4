"""B200GaussianPolicy -- drop-in for ``GaussianPolicy`` (mirl/policies/gaussian_policy.py:8-67) over
``MeanLogstdGaussian`` (mirl/torch_modules/gaussian.py:170-204) for the rollout loop.

``MBPOCollector.imagine`` calls ``policy.action(o)`` once per horizon step (mbpo_collector.py:48): in the
reference that is an eager 17 -> 256 -> 256 -> 12 MLP, ``chunk``, ``tanh`` bound, ``exp``, ``Normal.rsample``,
``tanh`` -- about 15 ATen kernels between two launches of the fused rollout step.  Here it is the layer GEMMs
of the ensemble kernels (a 1-member "ensemble") plus ONE head kernel (``mb_policy_action``): log-std bound,
reparameterised sample from a caller-visible noise tensor, tanh squashing and the log-probability.

Same constructor arguments, ``action()`` keywords, return values and state-dict keys
(``module.net.{2l}.weight`` ``[in,out]``, ``module.net.{2l}.bias`` ``[1,out]``) as the reference, so reference
checkpoints load unchanged.  ``action()`` is forward-only (model rollouts, exploration, target actions); the SAC actor
update -- which differentiates through this policy and the critics -- is ``B200ActorTrainer`` below.  The noise is drawn like the
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


class B200ActorTrainer(object):
    """Actor half of ``SACAgent.train_from_torch_batch`` (mirl/agents/sac_agent.py:170-195) on the device:
    ``policy.action(obs, reparameterize=True, return_log_prob=True)``, ``compute_alpha_loss`` + the Adam step on
    ``log_alpha`` (:88-100), ``compute_policy_loss`` (:102-118: ``alpha * log_prob.mean() - qf.value(obs, new_action,
    **current_v_pi_kwargs).mean()`` with the alpha AFTER its step) and the Adam step on the policy, as one chain of
    launches (``mb_actor_update``): tensor-core layer GEMMs forward and backward through the policy and (input
    gradients only) through the critics, hand-written sample / log-prob / reduction kernels.  The critics are read,
    not updated.

    ``policy`` is a ``B200GaussianPolicy``, ``qf`` a ``B200EnsembleQValue`` or ``B200TQCQValue``; the other
    arguments are the ``SACAgent`` ones (sac_agent.py:17-37).  ``current_v_pi_kwargs``: ``mode`` min / mean / max,
    ``sample_number`` (None = all members; fewer: the batch-wise draw of ensemble_value.py:84-86 -- the per-row draw
    is not covered); for TQC ``number_drop`` must be 0 (tqc_agent.py:26-28)."""

    def __init__(self, policy, qf, policy_lr=3e-4, alpha_if_not_automatic=1e-2, use_automatic_entropy_tuning=True,
                 init_log_alpha=0, target_entropy=None, current_v_pi_kwargs=None, optimizer_class="Adam",
                 optimizer_kwargs={}):
        import numpy as np
        if optimizer_class not in ("Adam", torch.optim.Adam):
            raise NotImplementedError("the actor update implements torch.optim.Adam only")
        extra = set(optimizer_kwargs) - {"betas", "eps"}
        if extra:
            raise NotImplementedError("Adam options %s" % sorted(extra))
        self.policy, self.qf = policy, qf
        self.policy_lr = policy_lr
        self.betas = tuple(optimizer_kwargs.get("betas", (0.9, 0.999)))
        self.eps = optimizer_kwargs.get("eps", 1e-8)
        self.use_automatic_entropy_tuning = use_automatic_entropy_tuning
        self.alpha_if_not_automatic = alpha_if_not_automatic
        self.target_entropy = (-float(np.prod(policy.action_shape)) if target_entropy is None else float(target_entropy))
        self.current_v_pi_kwargs = dict(current_v_pi_kwargs or {})
        dev = policy.device
        self.exp_avg = torch.zeros_like(policy.net.flat.data)
        self.exp_avg_sq = torch.zeros_like(policy.net.flat.data)
        # [log_alpha, exp_avg, exp_avg_sq]
        self.alpha_state = torch.tensor([float(init_log_alpha), 0.0, 0.0], device=dev)
        self.num_updates = 0
        self._ws = None
        self._alpha_host = None

    @property
    def log_alpha(self):
        return self.alpha_state[0:1]

    def set_log_alpha(self, value):
        """Restore ``log_alpha`` (e.g. from a checkpoint); its Adam moments are kept."""
        self.alpha_state[0] = float(value)
        self._alpha_host = None

    def get_alpha(self):
        """``SACAgent._get_alpha`` (sac_agent.py:66-72)."""
        if not self.use_automatic_entropy_tuning:
            return self.alpha_if_not_automatic
        # log_alpha only changes in update_async: one device read per actor update, not one per critic update
        # (REDQ runs 20 critic updates, each needing alpha for its TD target, per actor update)
        if self._alpha_host is None:
            self._alpha_host = float(self.alpha_state[0].exp().item())
        return self._alpha_host

    def _reduction(self):
        import numpy as np
        from .values import B200TQCQValue
        kw, E = self.current_v_pi_kwargs, self.qf.ensemble_size
        if isinstance(self.qf, B200TQCQValue):
            if kw.get("number_drop", 0) not in (0, None) or kw.get("percentage_drop") not in (0, None):
                raise NotImplementedError("the actor update pools all atoms (number_drop 0, tqc_agent.py:26-28)")
            return _lib.REDUCE["mean"], None
        mode = kw.get("mode", "min")
        if mode not in _lib.REDUCE:
            raise NotImplementedError("mode %r" % (mode,))
        n = kw.get("sample_number", 2)
        n = E if n is None else n
        members = None
        if n != E:
            if not kw.get("batchwise_sample", False):
                raise NotImplementedError("per-row member draws (sample_from_ensemble) inside the actor update")
            members = np.random.choice(E, n, replace=False).astype(np.int32)       # ensemble_value.py:85
        return _lib.REDUCE[mode], members

    def update_async(self, obs, noise=None):
        """One actor (+ alpha) update; returns the device tensor ``info [8]`` (see ``mb_actor_update``) without
        synchronising.  ``noise [B, A]`` replays a recorded ``rsample()`` draw."""
        pol, qf, L = self.policy, self.qf, _lib.lib()
        dev = pol.device
        obs = _lib.f32(obs, dev)
        assert obs.dim() == 2
        B, O, A = obs.shape[0], pol.observation_shape[0], pol.action_shape[0]
        noise = torch.randn(B, A, device=dev) if noise is None else _lib.f32(noise, dev)
        assert noise.shape == (B, A)
        for name in ("exp_avg", "exp_avg_sq", "alpha_state"):
            if getattr(self, name).device != dev:
                setattr(self, name, getattr(self, name).to(dev))
        pd, cd = ctypes.byref(pol.net.desc), ctypes.byref(qf.module.desc)
        need = L.mb_actor_workspace_bytes(pd, cd, B)
        if need == 0:
            raise NotImplementedError("actor update: %s" % L.mb_last_error().decode())
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        mode, members = self._reduction()
        marr = None if members is None else (ctypes.c_int32 * len(members))(*[int(m) for m in members])
        cfg = _lib.PolicyCfg(float(pol.mean_scale), float(pol.std_scale), float(pol.min_std),
                             float(pol.logstd_extra_bias), _lib.LOGSTD_BOUND[pol.bound_mode], int(bool(pol.squashed)))
        info = torch.zeros(8, device=dev)
        auto = self.use_automatic_entropy_tuning
        st = self.alpha_state
        with torch.cuda.device(dev):
            _lib.check(L.mb_actor_update(
                pd, _lib.ptr(pol.net.flat.data), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.num_updates,
                float(self.policy_lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), ctypes.byref(cfg),
                cd, _lib.ptr(qf.module.flat.data), mode, marr, 0 if members is None else len(members),
                _lib.ptr(obs), _lib.ptr(noise), O, A, B,
                _lib.ptr(st[0:1]) if auto else None, _lib.ptr(st[1:2]) if auto else None, _lib.ptr(st[2:3]) if auto else None,
                float(self.target_entropy), float(self.alpha_if_not_automatic), _lib.ptr(info), _lib.ptr(self._ws),
                self._ws.numel(), _lib.stream(dev)))
        self.num_updates += 1
        self._alpha_host = None
        pol.net.mark_weights_changed()
        return info

    def update(self, obs, noise=None):
        """-> the ``train_info`` entries of the reference (sac_agent.py:97-99,147-149), synchronising like it."""
        v = self.update_async(obs, noise).tolist()
        if self.use_automatic_entropy_tuning:
            self._alpha_host = v[5]                       # exp(log_alpha) after its step, as the kernel computed it
        out = {"policy/loss": v[0], "policy/q_pi": v[1], "policy/entropy": v[2]}
        if self.use_automatic_entropy_tuning:
            out.update({"alpha/value": v[3], "alpha/loss": v[4]})
        return out
