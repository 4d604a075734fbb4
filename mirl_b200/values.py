"""Critic-ensemble values with the reference's ``value()`` API, computed by CUDA kernels.

``B200EnsembleQValue`` mirrors ``EnsembleQValue`` (mirl/values/ensemble_value.py:36-120, REDQ's
random-subset target) and ``B200TQCQValue`` mirrors ``TQCQValue``
(mirl/values/distributional_value.py:10-102, TQC's truncated-quantile target).  Host RNG draws
(``np.random.choice`` for the member subset) are made exactly like the reference, in the same
order, so a seeded run consumes the global NumPy stream identically.

``value()`` is forward-only (the target network ``qf_target`` is evaluated under
``torch.no_grad()``, sac_agent.py:73-86, tqc_agent.py:41-57).  ``parameters()`` yields per-member
views in the reference's parameter order, so ``ptu.soft_update_from_to(qf, qf_target, tau)``
(torch_modules/utils.py:151-155) works with a reference ``qf`` as the source.

``B200CriticTrainer`` is the critic half of ``SACAgent.train_from_torch_batch`` (sac_agent.py:160-166) for
a ``B200EnsembleQValue`` pair: ``compute_qf_loss`` (ddpg_agent.py:138-149) + ``qf_optimizer.step()`` +
``_update_target`` (ddpg_agent.py:121-125) as ONE persistent kernel launch --
REDQ runs 20 of these per environment step.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .compat import QValue, snapshot_dir
from .ensemble import PackedEnsembleMLP


def sample_from_ensemble_indices(ensemble_size, batch_size, sample_number, replace=True):
    """Index draws of ``sample_from_ensemble`` (ensemble_value.py:10-34), ``[n, B]``."""
    if replace is False:
        assert sample_number <= ensemble_size
    if replace or sample_number == 1:
        idx = np.random.choice(ensemble_size, batch_size * sample_number)
        return idx.reshape(sample_number, batch_size)
    idx = np.stack([np.random.choice(ensemble_size - i, (batch_size,)) for i in range(sample_number)])
    for i in range(sample_number - 1, 0, -1):
        for j in range(i, sample_number):
            idx[j] = (idx[j] + idx[i - 1] + 1) % (ensemble_size + 1 - i)
    return idx


class _CriticBase(nn.Module, QValue):
    def _init(self, env, ensemble_size, out, value_name, mlp_kwargs):
        nn.Module.__init__(self)
        QValue.__init__(self, env)
        assert ensemble_size is not None
        self.ensemble_size = ensemble_size
        self.module = PackedEnsembleMLP(self._get_feature_size(), out, mlp_kwargs.pop("hidden_layers"),
                                        ensemble_size, module_name=value_name, **mlp_kwargs)

    def _get_feature_size(self):
        return self.observation_shape[0] + self.action_shape[0]

    @property
    def device(self):
        return self.module.flat.device

    def _workspace(self, B):
        need = _lib.lib().mb_critic_workspace_bytes(ctypes.byref(self.module.desc), B)
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need or ws.device != self.device:
            ws = self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def parameters(self, recurse=True):
        """Per-member views in the reference's order (``net.{2l}.weight_net{e}``,
        ``net.{2l}.bias_net{e}`` per layer, per member) -- see the module docstring."""
        mod = self.module
        for l in range(mod.num_fc):
            for e in range(mod.ensemble_size):
                yield mod.weight(l)[e]
                yield mod.bias(l)[e]

    def soft_update_from(self, source, tau):
        """``ptu.soft_update_from_to(source, self, tau)`` in one fused lerp when ``source`` is
        also packed (the reference loops over 6*E tensors)."""
        if isinstance(source, _CriticBase):
            self.module.flat.data.lerp_(source.module.flat.data, tau)
        else:
            for tp, p in zip(self.parameters(), source.parameters()):
                tp.copy_(tp * (1.0 - tau) + p.data * tau)

    def _rows(self, obs, action):
        """Flatten CQL-style leading batch dims (ensemble_value.py:59-63): ``[..., B, F]``."""
        dev = self.device
        obs, action = _lib.f32(obs, dev), _lib.f32(action, dev)
        assert obs.dim() == action.dim() and obs.dim() >= 2
        lead = tuple(obs.shape[:-1])
        return obs.reshape(-1, obs.shape[-1]), action.reshape(-1, action.shape[-1]), lead

    @_lib.on_device
    def _ensemble(self, obs2, act2):
        E, N = self.ensemble_size, self.module.output_size
        B = obs2.shape[0]
        out = torch.empty(E, B, N, device=self.device)
        mod = self.module
        ws = self._workspace(B)
        _lib.check(_lib.lib().mb_critic_forward(
            ctypes.byref(mod.desc), _lib.ptr(mod.flat.data), _lib.ptr(obs2), _lib.ptr(act2),
            self.observation_shape[0], self.action_shape[0], B, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
            _lib.stream()))
        return out

    def _ensemble_shaped(self, obs2, act2, lead):
        out = self._ensemble(obs2, act2)                        # [E, rows, N]
        if len(lead) == 1:
            return out
        # reference: features [..., 1, B, F] -> MLP -> [..., E, B, N]
        out = out.reshape((self.ensemble_size,) + lead + (out.shape[-1],))
        return out.movedim(0, -3).contiguous()

    def save(self, save_dir=None):
        self.module.save(save_dir or snapshot_dir())

    def load(self, load_dir=None):
        self.module.load(load_dir or snapshot_dir())

    def get_snapshot(self):
        return self.module.get_snapshot()


class B200EnsembleQValue(_CriticBase):
    def __init__(self, env, ensemble_size=2, value_name="ensemble_q_value", **mlp_kwargs):
        self._init(env, ensemble_size, 1, value_name, mlp_kwargs)

    def _draw_subset(self, rows, lead, sample_number, batchwise_sample):
        """Host RNG of ensemble_value.py:80-90; returns (members list | None, row table | None)."""
        E = self.ensemble_size
        if sample_number is None:
            sample_number = E
        if E == sample_number:
            return None, None, sample_number
        if batchwise_sample:
            return np.random.choice(E, sample_number, replace=False), None, sample_number
        assert len(lead) == 1, "per-row sampling needs a 3-D ensemble tensor (ensemble_value.py:16)"
        return None, sample_from_ensemble_indices(E, rows, sample_number, replace=False), sample_number

    def _target(self, obs2, act2, members, row_table, n_sub, mode, td=None):
        mod, dev = self.module, self.device
        B = obs2.shape[0]
        value = torch.empty(B, 1, device=dev)
        ws = self._workspace(B)
        mem = rows = None
        if members is not None:
            mem = (ctypes.c_int32 * len(members))(*[int(i) for i in members])
        if row_table is not None:
            rows = torch.from_numpy(np.ascontiguousarray(row_table.astype(np.uint8))).to(dev)
        r = t = lp = None
        alpha = disc = 0.0
        if td is not None:
            r, t, lp = (_lib.f32(x, dev).reshape(B, 1) for x in (td["rewards"], td["terminals"], td["log_prob"]))
            alpha, disc = float(td["alpha"]), float(td["discount"])
        _lib.check(_lib.lib().mb_critic_redq_target(
            ctypes.byref(mod.desc), _lib.ptr(mod.flat.data), _lib.ptr(obs2), _lib.ptr(act2),
            self.observation_shape[0], self.action_shape[0], B, mem, _lib.ptr(rows), int(n_sub),
            _lib.REDUCE[mode], int(td is not None), _lib.ptr(r), _lib.ptr(t), _lib.ptr(lp), alpha, disc,
            _lib.ptr(value), _lib.ptr(ws), ws.numel(), _lib.stream()))
        return value

    @_lib.on_device
    def value(self, obs, action, sample_number=2, batchwise_sample=False, mode="min",
              return_ensemble=False, only_return_ensemble=False, _td=None):
        """ensemble_value.py:65-107; only the drawn members are evaluated."""
        obs2, act2, lead = self._rows(obs, action)
        if only_return_ensemble:
            return None, {"ensemble_value": self._ensemble_shaped(obs2, act2, lead)}
        members, row_table, n_sub = self._draw_subset(obs2.shape[0], lead, sample_number,
                                                      batchwise_sample)
        if mode == "sample":
            index = np.random.randint(n_sub)                        # ensemble_value.py:98-99
            if members is not None:
                members = [members[index]]
            elif row_table is not None:
                row_table = row_table[index:index + 1]
            else:
                members = [index]
            n_sub, mode = 1, "min"
        elif mode not in _lib.REDUCE:
            raise NotImplementedError
        value = self._target(obs2, act2, members, row_table, n_sub, mode, _td).reshape(lead + (1,))
        info = {}
        if return_ensemble:
            info["ensemble_value"] = self._ensemble_shaped(obs2, act2, lead)
        return value, info

    @_lib.on_device
    def q_target(self, next_obs, next_action, rewards, terminals, log_prob, alpha, discount=0.99,
                 **v_pi_kwargs):
        """``SACAgent.compute_q_target`` arithmetic (sac_agent.py:80-85) fused into the value
        kernel: ``r + (1-d)*discount*(value(s',a') - alpha*log_prob)``."""
        td = dict(rewards=rewards, terminals=terminals, log_prob=log_prob, alpha=alpha, discount=discount)
        return self.value(next_obs, next_action, _td=td, **v_pi_kwargs)[0]


class B200TQCQValue(_CriticBase):
    def __init__(self, env, ensemble_size=5, number_head=25, value_name="tqc_q_value", **mlp_kwargs):
        self.number_head = number_head
        self._init(env, ensemble_size, number_head, value_name, mlp_kwargs)

    def _n_drop(self, number_drop, percentage_drop):
        n_atom = self.ensemble_size * self.number_head
        if number_drop is None:                                     # distributional_value.py:39-44
            number_drop = 0 if percentage_drop is None else round(n_atom * percentage_drop / 100)
        if number_drop < 0:
            raise NotImplementedError
        return number_drop, n_atom

    def _atoms(self, obs2, act2, n_drop, n_atom, td=None):
        mod, dev = self.module, self.device
        B = obs2.shape[0]
        value = torch.empty(B, 1, device=dev)
        atoms = torch.empty(B, n_atom - n_drop, device=dev)
        ws = self._workspace(B)
        r = t = lp = None
        alpha = disc = 0.0
        if td is not None:
            r, t, lp = (_lib.f32(x, dev).reshape(B, 1) for x in (td["rewards"], td["terminals"], td["log_prob"]))
            alpha, disc = float(td["alpha"]), float(td["discount"])
        _lib.check(_lib.lib().mb_critic_tqc_target(
            ctypes.byref(mod.desc), _lib.ptr(mod.flat.data), _lib.ptr(obs2), _lib.ptr(act2),
            self.observation_shape[0], self.action_shape[0], B, int(n_drop), int(td is not None),
            _lib.ptr(r), _lib.ptr(t), _lib.ptr(lp), alpha, disc, _lib.ptr(value), _lib.ptr(atoms),
            _lib.ptr(ws), ws.numel(), _lib.stream()))
        return value, atoms

    @_lib.on_device
    def value(self, obs, action, number_drop=None, percentage_drop=None, return_atoms=False,
              return_ensemble=False, only_return_ensemble=False):
        """distributional_value.py:59-93."""
        obs2, act2, lead = self._rows(obs, action)
        if only_return_ensemble:
            return None, {"ensemble_atoms": self._ensemble_shaped(obs2, act2, lead)}
        n_drop, n_atom = self._n_drop(number_drop, percentage_drop)
        value, atoms = self._atoms(obs2, act2, n_drop, n_atom)
        info = {}
        if return_atoms:
            info["value_atoms"] = atoms.reshape(lead + (atoms.shape[-1],))
        if return_ensemble:
            ens = self._ensemble_shaped(obs2, act2, lead)           # [..., E, B, Q]
            ens = ens.transpose(-3, -2)
            info["ensemble_atoms"] = ens.reshape(ens.shape[:-2] + (-1,))
        return value.reshape(lead + (1,)), info

    @_lib.on_device
    def atoms_target(self, next_obs, next_action, rewards, terminals, log_prob, alpha,
                     discount=0.99, number_drop=None, percentage_drop=None):
        """``TQCAgent.compute_q_target`` (tqc_agent.py:41-57) fused: truncated sorted atoms of
        the target critics through ``r + (1-d)*discount*(atoms - alpha*log_prob)``."""
        obs2, act2, lead = self._rows(next_obs, next_action)
        n_drop, n_atom = self._n_drop(number_drop, percentage_drop)
        td = dict(rewards=rewards, terminals=terminals, log_prob=log_prob, alpha=alpha, discount=discount)
        return self._atoms(obs2, act2, n_drop, n_atom, td)[1]


class B200CriticTrainer(object):
    """Fused critic update for REDQ / SAC ensembles: forward over all E critics, ``F.mse_loss`` against
    the expanded target (optional ``clip_error``), backward, ``torch.optim.Adam`` and the soft target
    update, in one launch (``mb_critic_train_step``; the reference issues ~300 ATen calls plus a Python loop
    over 6*E parameter tensors, torch_modules/utils.py:151-155).  Critics wider than 256 run the same update
    as a chain of tensor-core layer GEMM launches.

    ``qf`` and ``qf_target`` are ``B200EnsembleQValue`` of the same shape; ``qf_lr``, ``soft_target_tau``,
    ``target_update_freq`` and ``clip_error`` are the ``DDPGAgent`` arguments (ddpg_agent.py:44-60).

    With ``B200TQCQValue`` critics the loss is TQC's ``quantile_huber_loss_f`` (tqc_agent.py:5-14,59-66) and
    ``q_target`` is ``atoms_target [B, n_target]`` (``mb_tqc_train_step``)."""

    def __init__(self, qf, qf_target=None, qf_lr=3e-4, soft_target_tau=5e-3, target_update_freq=1,
                 clip_error=None, optimizer_class="Adam", optimizer_kwargs={}):
        if optimizer_class not in ("Adam", torch.optim.Adam):
            raise NotImplementedError("the fused critic update implements torch.optim.Adam only")
        extra = set(optimizer_kwargs) - {"betas", "eps"}
        if extra:
            raise NotImplementedError("Adam options %s" % sorted(extra))
        self.qf, self.qf_target = qf, qf_target
        if qf_target is not None:
            assert tuple(qf.module.flat.shape) == tuple(qf_target.module.flat.shape)
        self.qf_lr = qf_lr
        self.betas = tuple(optimizer_kwargs.get("betas", (0.9, 0.999)))
        self.eps = optimizer_kwargs.get("eps", 1e-8)
        self.soft_target_tau = soft_target_tau
        self.target_update_freq = target_update_freq
        self.clip_error = clip_error
        self.exp_avg = torch.zeros_like(qf.module.flat.data)
        self.exp_avg_sq = torch.zeros_like(qf.module.flat.data)
        self.num_train_steps = 0
        self._ws = None

    def supported(self):
        return _lib.lib().mb_critic_train_workspace_bytes(ctypes.byref(self.qf.module.desc), 1) > 0

    def update_async(self, obs, actions, q_target):
        """One critic update; returns device tensors ``(qf_loss [], ensemble_loss [E])`` without
        synchronising.  ``q_target`` is ``[B, out]`` (``compute_q_target``'s result, detached)."""
        qf, L = self.qf, _lib.lib()
        dev = qf.device
        mod = qf.module
        obs, actions, q_target = (_lib.f32(t, dev) for t in (obs, actions, q_target))
        B = obs.shape[0]
        quantile = isinstance(qf, B200TQCQValue)
        assert obs.dim() == 2 and q_target.dim() == 2 and q_target.shape[0] == B
        assert quantile or q_target.shape[1] == mod.output_size
        if quantile and self.clip_error:
            raise NotImplementedError("clip_error applies to the MSE critic loss only (ddpg_agent.py:142-146)")
        if self.exp_avg.device != dev:
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(dev), self.exp_avg_sq.to(dev)
        need = L.mb_critic_train_workspace_bytes(ctypes.byref(mod.desc), B)
        if need == 0:
            raise NotImplementedError("critic shape not covered: %s" % _lib.lib().mb_last_error().decode())
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        el = torch.zeros(mod.ensemble_size, device=dev)
        terms = torch.zeros(3, device=dev)
        tgt = None
        if self.qf_target is not None and self.num_train_steps % self.target_update_freq == 0:
            tgt = self.qf_target.module.flat.data
        head = (ctypes.byref(mod.desc), _lib.ptr(mod.flat.data), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                self.num_train_steps, float(self.qf_lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                _lib.ptr(obs), _lib.ptr(actions), qf.observation_shape[0], qf.action_shape[0], _lib.ptr(q_target))
        tail = (_lib.ptr(tgt), float(self.soft_target_tau), _lib.ptr(el), _lib.ptr(terms), _lib.ptr(self._ws),
                self._ws.numel(), _lib.stream(dev))
        with torch.cuda.device(dev):
            if quantile:
                _lib.check(L.mb_tqc_train_step(*head, int(q_target.shape[1]), B, *tail))
            else:
                _lib.check(L.mb_critic_train_step(*head, B, float(self.clip_error or 0.0), *tail))
        self.num_train_steps += 1
        mod.mark_weights_changed()
        if tgt is not None:
            self.qf_target.module.mark_weights_changed()
        return terms[0], el

    def update(self, obs, actions, q_target):
        """-> ``{'qf/loss': float}`` (ddpg_agent.py:155), synchronising like the reference."""
        loss, _ = self.update_async(obs, actions, q_target)
        return {"qf/loss": float(loss.item())}
