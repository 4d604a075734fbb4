"""CPU oracle for the critic-ensemble targets (REDQ / TQC).  TEST INFRASTRUCTURE ONLY.

Same rules as ``oracle/pe_oracle.py``: a restatement of the reference arithmetic, pinned by the
fixtures under ``tests/golden/`` that were generated from the unmodified reference; imported
only by tests, ``smoke()`` and the CPU-baseline legs of ``bench.py``.
"""
import numpy as np
import torch

from .pe_oracle import mlp_forward


def subset_indices_batchwise(ensemble_size, sample_number, rng=np.random):
    """EnsembleQValue.value batch-wise branch, mirl/values/ensemble_value.py:84-86."""
    return rng.choice(ensemble_size, sample_number, replace=False)


def subset_indices_rowwise(ensemble_size, batch_size, sample_number, rng=np.random):
    """sample_from_ensemble(replace=False), mirl/values/ensemble_value.py:20-31.
    Returns ``[sample_number, B]`` member indices (distinct within a column)."""
    if sample_number == 1:
        return rng.choice(ensemble_size, batch_size * sample_number).reshape(1, batch_size)
    indices = [rng.choice(ensemble_size - i, (batch_size,)) for i in range(sample_number)]
    indices = np.stack(indices)
    for i in range(sample_number - 1, 0, -1):
        for j in range(i, sample_number):
            indices[j] = (indices[j] + indices[i - 1] + 1) % (ensemble_size + 1 - i)
    return indices


def reduce_members(sampled, mode):
    """ensemble_value.py:91-96 (min / mean / max over the member dim, -3)."""
    if mode == "min":
        return torch.min(sampled, dim=-3)[0]
    if mode == "mean":
        return torch.mean(sampled, dim=-3)
    if mode == "max":
        return torch.max(sampled, dim=-3)[0]
    raise NotImplementedError(mode)


def ensemble_q(W, b, obs, act, activation="relu"):
    """EnsembleQValue forward, ensemble_value.py:59-63,75-77: ``cat[obs,act]`` -> MLP ->
    ``[E,B,1]``.  Evaluates every member, like the reference."""
    return mlp_forward(W, b, torch.cat([obs, act], dim=-1), activation)


def redq_value(W, b, obs, act, member_idx=None, row_idx=None, mode="min", activation="relu"):
    """EnsembleQValue.value for a drawn subset.  ``member_idx [n]`` = batch-wise draw
    (ensemble_value.py:84-86); ``row_idx [n,B]`` = per-row draw (:10-34); neither = all
    members (:89-90)."""
    q = ensemble_q(W, b, obs, act, activation)
    if member_idx is not None:
        q = q[..., torch.as_tensor(np.asarray(member_idx), dtype=torch.long), :, :]
    elif row_idx is not None:
        row_idx = torch.as_tensor(np.asarray(row_idx), dtype=torch.long)
        ar = torch.arange(q.shape[1]).expand_as(row_idx)
        q = q[row_idx, ar]
    return reduce_members(q, mode)


def td_target(value, rewards, terminals, log_prob, alpha, discount=0.99):
    """SACAgent.compute_q_target arithmetic, mirl/agents/sac_agent.py:80-85 (also
    tqc_agent.py:54-56 when ``value`` holds atoms ``[B,K]``)."""
    return rewards + (1.0 - terminals) * discount * (value - alpha * log_prob)


def tqc_atoms(W, b, obs, act, number_drop=None, percentage_drop=None, activation="relu"):
    """TQCQValue.value(return_atoms=True), mirl/values/distributional_value.py:59-93 and
    ``_drop_atoms`` :32-57.  Returns ``value [B,1]`` and kept atoms ``[B, E*Q - drop]``."""
    atoms = mlp_forward(W, b, torch.cat([obs, act], dim=-1), activation)   # [E,B,Q]
    atoms = atoms.transpose(-3, -2)
    atoms = atoms.reshape(atoms.shape[:-2] + (-1,))                          # [B,E*Q]
    n_atom = atoms.shape[-1]
    if number_drop is None:
        number_drop = 0 if percentage_drop is None else round(n_atom * percentage_drop / 100)
    if number_drop > 0:
        atoms, _ = torch.sort(atoms, dim=-1)
        atoms = atoms[..., :-number_drop]
    return torch.mean(atoms, dim=-1, keepdim=True), atoms


def quantile_huber(samples, atoms):
    """quantile_huber_loss_f, mirl/agents/tqc_agent.py:5-14: ``samples [B,T]`` target atoms, ``atoms [E,B,Q]``.
    Returns the loss and its hand-written gradient with respect to ``atoms``."""
    delta = samples[None, :, :, None] - atoms[..., None, :]               # [E,B,T,Q]
    ad = delta.abs()
    huber = torch.where(ad > 1, ad - 0.5, 0.5 * delta * delta)
    Q = atoms.shape[-1]
    tau = torch.arange(Q, dtype=atoms.dtype) / Q + 0.5 / Q
    wgt = (tau - (delta < 0).to(atoms.dtype)).abs()
    loss = (wgt * huber).mean()
    grad = -(wgt * delta.clamp(-1.0, 1.0)).sum(2) / delta.numel()
    return loss, grad


def critic_update(W, b, opt, target_W, target_b, obs, act, q_target, lr=3e-4, tau=5e-3, clip_error=None,
                  activation="relu", head="mse"):
    """Critic half of SACAgent.train_from_torch_batch (mirl/agents/sac_agent.py:160-166):
    DDPGAgent.compute_qf_loss (ddpg_agent.py:138-149: F.mse_loss of the [E,B,out] ensemble prediction against
    the expanded target, optional clip_error), hand-written backward, torch.optim.Adam (ddpg_agent.py:82-86)
    and ptu.soft_update_from_to (torch_modules/utils.py:151-155).  Mutates W, b, opt and the target lists;
    ``opt = {"step": 0, "m": [...], "v": [...]}`` over ``W + b``.  Returns the loss."""
    from oracle.pe_oracle import adam_update
    assert activation == "relu"
    L = len(W)
    x = torch.cat([obs, act], -1).unsqueeze(0).expand(W[0].shape[0], -1, -1)
    hs, zs = [x], []
    for l in range(L):
        z = hs[-1].matmul(W[l]) + b[l]
        zs.append(z)
        if l < L - 1:
            hs.append(torch.relu(z))
    q = zs[-1]
    if head == "quantile":                 # TQCAgent.compute_qf_loss, tqc_agent.py:59-66; q_target = atoms [B,T]
        loss, g = quantile_huber(q_target, q)
    else:
        diff = q - q_target.unsqueeze(0)
        if clip_error is not None and clip_error > 0:
            diff = torch.clamp(diff, -clip_error, clip_error)
        loss = (diff * diff).mean()
        g = 2.0 * diff / diff.numel()
    gW, gb = [None] * L, [None] * L
    for l in range(L - 1, -1, -1):
        gW[l] = hs[l].transpose(1, 2).matmul(g)
        gb[l] = g.sum(1, keepdim=True)
        if l > 0:
            g = g.matmul(W[l].transpose(1, 2)) * (zs[l - 1] > 0).float()
    opt["step"] += 1
    for t, gr, m, v in zip(list(W) + list(b), gW + gb, opt["m"], opt["v"]):
        adam_update(t, gr, m, v, opt["step"], lr=lr)
    for tp, p in zip(list(target_W) + list(target_b), list(W) + list(b)):
        tp.copy_(tp * (1.0 - tau) + p * tau)
    return loss


def policy_action(W, b, obs, eps=None, squashed=True, min_std=1e-4, mean_scale=1.0, std_scale=1.0,
                  logstd_extra_bias=0.0, activation="relu"):
    """GaussianPolicy.action -> Gaussian.forward over MeanLogstdGaussian.get_mean_std with bound_mode 'tanh'
    (mirl/policies/gaussian_policy.py:35-47, mirl/torch_modules/gaussian.py:17-30,55-97,199-204).  ``W``/``b`` are
    1-member lists ``[1,in,out]`` / ``[1,1,out]``; ``eps`` is the Normal draw (None = deterministic).
    Returns action, log_prob [B,1], mean, std."""
    out = mlp_forward(W, b, obs, activation)[0]
    mean, ls = torch.chunk(out, 2, -1)
    ls = (torch.tanh(ls + logstd_extra_bias) + 1) * 6.0 - 10.0          # (LOG_STD_MAX - LOG_STD_MIN) / 2 = 6
    mean, std = mean * mean_scale, torch.exp(ls) * std_scale + min_std
    pre = mean if eps is None else mean + std * eps
    lp = -((pre - mean) ** 2) / (2 * std ** 2) - torch.log(std) - 0.5 * np.log(2 * np.pi)
    if squashed:
        lp = lp - (2 * np.log(2) - torch.nn.functional.softplus(2 * pre) - torch.nn.functional.softplus(-2 * pre))
    return (torch.tanh(pre) if squashed else pre), lp.sum(-1, keepdim=True), mean, std


def actor_update(pW, pb, popt, log_alpha, aopt, cW, cb, obs, eps, lr=3e-4, target_entropy=-1.0, mode="mean",
                 member_idx=None, alpha_fixed=None, squashed=True, activation="relu"):
    """Actor half of SACAgent.train_from_torch_batch (mirl/agents/sac_agent.py:170-195): policy.action with
    ``reparameterize=True, return_log_prob=True`` (gaussian.py:55-97; ``eps`` is the rsample draw),
    compute_alpha_loss (:88-100) + Adam on ``log_alpha``, compute_policy_loss (:102-118) with the alpha AFTER its
    step, ``qf.value(obs, new_action, mode=..., sample_number/batchwise index=member_idx)`` (ensemble_value.py:65-107;
    several outputs per member: TQCQValue.value with number_drop 0 = mean of all atoms, distributional_value.py:51-55),
    Adam on the policy.  Gradients by torch autograd on this restatement.  Mutates ``pW, pb, popt`` (policy
    parameters ``[1,in,out]`` / ``[1,1,out]`` and ``{"step","m","v"}``) and ``log_alpha`` / ``aopt`` (a 1-element
    tensor and its state; ``log_alpha=None``: fixed ``alpha_fixed``).  Returns the info dict of the reference
    (policy/loss, policy/q_pi, policy/entropy, alpha/value, alpha/loss)."""
    from oracle.pe_oracle import adam_update
    leaves = [t.clone().requires_grad_(True) for t in list(pW) + list(pb)]
    n = len(pW)
    action, lp, _, _ = policy_action(leaves[:n], leaves[n:], obs, eps, squashed=squashed, activation=activation)
    info = {}
    if log_alpha is not None:
        res = float(lp.detach().mean()) + target_entropy
        alpha = float(log_alpha.exp())
        info["alpha/value"], info["alpha/loss"] = alpha, -res * alpha
        aopt["step"] += 1
        adam_update(log_alpha, torch.tensor([-res * alpha]), aopt["m"][0], aopt["v"][0], aopt["step"], lr=lr)
        alpha = float(log_alpha.exp())
    else:
        alpha = alpha_fixed
    q = mlp_forward(cW, cb, torch.cat([obs, action], dim=-1), activation)           # [E,B,Q]
    if member_idx is not None:
        q = q[torch.as_tensor(np.asarray(member_idx), dtype=torch.long)]
    if q.shape[-1] > 1:                                                           # TQC: pooled atoms, no drop
        value = q.transpose(0, 1).reshape(q.shape[1], -1).mean(-1, keepdim=True)
    else:
        value = reduce_members(q, mode)
    entropy, q_pi = -lp.mean(), value.mean()
    loss = -alpha * entropy - q_pi
    grads = torch.autograd.grad(loss, leaves)
    popt["step"] += 1
    for t, g, m, v in zip(list(pW) + list(pb), grads, popt["m"], popt["v"]):
        adam_update(t, g, m, v, popt["step"], lr=lr)
    info.update({"policy/loss": float(loss.detach()), "policy/q_pi": float(q_pi.detach()),
                 "policy/entropy": float(entropy.detach())})
    return info


def sac_train_step(state, batch, step, noise_fn, discount=0.99, lr=3e-4, tau=5e-3, target_entropy=-1.0,
                   policy_update_freq=20, next_sample_number=2, current_mode="mean"):
    """SACAgent.train_from_torch_batch (mirl/agents/sac_agent.py:152-204) with REDQ's settings, composed from the
    functions above, consuming NumPy's global generator exactly like the reference: the batch-wise member pair of
    the TD target (ensemble_value.py:84-86) and the per-row sample that ``compute_qf_loss`` draws and discards
    (``qf.value(..., return_ensemble=True)`` with the default sample_number 2, ddpg_agent.py:139 ->
    ensemble_value.py:87-88).  ``state``: dict with pW, pb, popt, cW, cb, copt, tW, tb, log_alpha, aopt;
    ``noise_fn(B, A)`` = the policy's standard-normal draw.  Returns the reference's train_info entries."""
    pW, pb, cW, cb, tW, tb = (state[k] for k in ("pW", "pb", "cW", "cb", "tW", "tb"))
    B, A = batch["actions"].shape
    E = cW[0].shape[0]
    alpha = float(state["log_alpha"].exp())
    next_action, lp, _, _ = policy_action(pW, pb, batch["next_observations"], noise_fn(B, A))
    idx = subset_indices_batchwise(E, next_sample_number)
    v = redq_value(tW, tb, batch["next_observations"], next_action, member_idx=idx, mode="min")
    q_target = td_target(v, batch["rewards"], batch["terminals"], lp, alpha, discount)
    if E != 2:
        subset_indices_rowwise(E, B, 2)                       # drawn and discarded by compute_qf_loss
    info = {"qf/loss": float(critic_update(cW, cb, state["copt"], tW, tb, batch["observations"], batch["actions"],
                                           q_target, lr=lr, tau=tau))}
    if step % policy_update_freq == 0:
        info.update(actor_update(pW, pb, state["popt"], state["log_alpha"], state["aopt"], cW, cb,
                                 batch["observations"], noise_fn(B, A), lr=lr, target_entropy=target_entropy,
                                 mode=current_mode))
    return info
