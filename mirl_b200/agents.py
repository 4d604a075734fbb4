"""B200SACUpdater -- ``SACAgent.train_from_torch_batch`` (mirl/agents/sac_agent.py:152-204) on the device.

The update step of SAC / REDQ / TQC-style agents, assembled from the classes of this package: the TD target
(``compute_q_target`` sac_agent.py:73-86 / tqc_agent.py:41-57: policy sample for the next state + the fused
critic-target kernel), the critic update (``B200CriticTrainer``: loss + Adam + soft target update), and every
``policy_update_freq`` steps the actor and entropy-coefficient update (``B200ActorTrainer``).  Same call order, same
RNG consumption (one standard-normal ``[B, A]`` draw per ``policy.action`` call, NumPy's global generator for the
critics' member subsets), same ``train_info`` keys as the reference.  The reference agent keeps its exploration,
logging and evaluation code; only the body of ``train_from_torch_batch`` is replaced:

    self.updater = mirl_b200.B200SACUpdater(self.policy, self.qf, self.qf_target, discount=self.discount, ...)
    def train_from_torch_batch(self, batch):
        self.train_info.update(self.updater.train_from_torch_batch(batch)); self.num_train_steps += 1
"""
import torch

from .policies import B200ActorTrainer
from .values import B200CriticTrainer, B200TQCQValue


class B200SACUpdater(object):
    """Arguments are those of ``SACAgent`` / ``DDPGAgent`` (sac_agent.py:17-37, ddpg_agent.py:40-60) /
    ``TQCAgent`` (tqc_agent.py:17-39).  ``noise_fn(shape) -> tensor`` replaces the device generator for the policy's
    draws (tests replay the reference's CPU stream with it).  ``sync_info=False`` leaves ``qf/loss`` on the device (a
    0-dim tensor), so that the critic updates between two actor updates run without a host read."""

    def __init__(self, policy, qf, qf_target, discount=0.99, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                 policy_update_freq=1, target_update_freq=1, clip_error=None, alpha_if_not_automatic=1e-2,
                 use_automatic_entropy_tuning=True, init_log_alpha=0, target_entropy=None, next_v_pi_kwargs=None,
                 current_v_pi_kwargs=None, optimizer_class="Adam", optimizer_kwargs={}, noise_fn=None, sync_info=True):
        self.policy, self.qf, self.qf_target = policy, qf, qf_target
        self.tqc = isinstance(qf, B200TQCQValue)
        if next_v_pi_kwargs is None:                      # tqc_agent.py:23-28 defaults
            next_v_pi_kwargs = {"percentage_drop": 8} if self.tqc else {}
        if current_v_pi_kwargs is None:
            current_v_pi_kwargs = {"number_drop": 0} if self.tqc else {}
        self.discount = discount
        self.policy_update_freq = policy_update_freq
        self.next_v_pi_kwargs = dict(next_v_pi_kwargs)
        self.noise_fn = noise_fn
        with torch.no_grad():                              # DDPGAgent.__init__: _update_target(1), ddpg_agent.py:68
            qf_target.module.flat.data.copy_(qf.module.flat.data)
        qf_target.module.mark_weights_changed()
        self.critic = B200CriticTrainer(qf, qf_target, qf_lr=qf_lr, soft_target_tau=soft_target_tau,
                                        target_update_freq=target_update_freq, clip_error=clip_error,
                                        optimizer_class=optimizer_class, optimizer_kwargs=optimizer_kwargs)
        self.actor = B200ActorTrainer(policy, qf, policy_lr=policy_lr, alpha_if_not_automatic=alpha_if_not_automatic,
                                      use_automatic_entropy_tuning=use_automatic_entropy_tuning,
                                      init_log_alpha=init_log_alpha, target_entropy=target_entropy,
                                      current_v_pi_kwargs=current_v_pi_kwargs, optimizer_class=optimizer_class,
                                      optimizer_kwargs=optimizer_kwargs)
        self.sync_info = sync_info
        self.num_train_steps = 0

    def _noise(self, B):
        if self.noise_fn is None:
            return None
        return self.noise_fn((B, self.policy.action_shape[0]))

    def compute_q_target(self, next_obs, rewards, terminals):
        """sac_agent.py:73-86 (critic ensembles) / tqc_agent.py:41-57 (atoms)."""
        alpha = self.actor.get_alpha()
        B = next_obs.shape[0]
        next_action, info = self.policy.action(next_obs, reparameterize=False, return_log_prob=True,
                                               noise=self._noise(B))
        if self.tqc:
            return self.qf_target.atoms_target(next_obs, next_action, rewards, terminals, info["log_prob"], alpha=alpha,
                                               discount=self.discount, **self.next_v_pi_kwargs)
        return self.qf_target.q_target(next_obs, next_action, rewards, terminals, info["log_prob"], alpha=alpha,
                                       discount=self.discount, **self.next_v_pi_kwargs)

    def train_from_torch_batch(self, batch):
        """One update; returns the ``train_info`` entries the reference produces for it."""
        obs, actions = batch["observations"], batch["actions"]
        q_target = self.compute_q_target(batch["next_observations"], batch["rewards"], batch["terminals"])
        if not self.tqc:
            # compute_qf_loss calls qf.value(obs, actions, return_ensemble=True) with the default sample_number = 2
            # (ddpg_agent.py:139): the reference draws -- and discards -- a per-row member sample there
            # (ensemble_value.py:80-90).  Consume the same numbers, so that a run stays on the reference's NumPy stream.
            self.qf._draw_subset(obs.shape[0], (1,), 2, False)
        if self.sync_info:
            info = dict(self.critic.update(obs, actions, q_target))
        else:                                      # no host read: the loss stays a 0-dim device tensor
            info = {"qf/loss": self.critic.update_async(obs, actions, q_target)[0]}
        if self.num_train_steps % self.policy_update_freq == 0:
            info.update(self.actor.update(obs, noise=self._noise(obs.shape[0])))
        self.num_train_steps += 1
        return info
