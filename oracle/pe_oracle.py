"""CPU oracle for the probabilistic-ensemble path.  TEST INFRASTRUCTURE ONLY.

A plain restatement (torch CPU fp32 tensors, explicit formulae, no autograd, no optimiser
object) of what the reference computes on this path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it -- as the checker or the timed CPU arm, never as the product path.  ``mirl_b200``
never imports ``oracle``.

Pinned: every function here is checked against outputs of the UNMODIFIED reference
(``tests/golden/*.npz`` made by ``tests/golden/gen_golden.py`` importing ``/root/reference``;
see ``tests/test_oracle_golden.py``).  The reference holds no golden vectors of its own
(SURVEY.md section 4), so those fixtures are the pin.

Citations are relative to the reference tree.
"""
import math

import numpy as np
import torch

LOG2PI = float(np.log(2 * np.pi))          # mirl/agents/mbpo/pe_model.py:13
NORM_EPS = 1e-6                            # mirl/processors/normalizer.py:7

DONE_KINDS = ("cheetah", "hopper", "walker2d", "mb_ant", "mb_humanoid", "inverted_pendulum")


# --------------------------------------------------------------------------- activations
def swish(x):
    """mirl/torch_modules/utils.py:125-130  x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def activation(name):
    if name == "swish":
        return swish
    if name == "relu":
        return torch.relu
    if name == "identity":
        return lambda x: x
    raise NotImplementedError(name)


# --------------------------------------------------------------------------- ensemble MLP
def mlp_forward(W, b, x, act="swish"):
    """MyEnsembleLinear + MLP forward.

    mirl/torch_modules/linear.py:143-149 (``x.matmul(w) + b`` with ``w [E,in,out]``,
    ``b [E,1,out]``; a 2-D input is broadcast over members) and
    mirl/torch_modules/mlp.py:14-38,101-109 (activation after every layer but the last).
    ``x`` is ``[B,in]`` (shared) or ``[E,B,in]`` (per member).  Returns ``[E,B,out]``.
    """
    f = activation(act)
    h = x
    while h.dim() < 3:
        h = h.unsqueeze(0)
    n = len(W)
    for l in range(n):
        h = h.matmul(W[l]) + b[l]
        if l < n - 1:
            h = f(h)
    return h


def softplus(x):
    """torch.nn.functional.softplus defaults (beta 1, threshold 20), written out."""
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(x)))


def mean_logstd(out, max_logstd, min_logstd):
    """PEModelModule._get_mean_logstd, mirl/agents/mbpo/pe_model.py:145-151."""
    d = out.shape[-1] // 2
    mean, ls = out[..., :d], out[..., d:]
    ls = max_logstd - softplus(max_logstd - ls)
    ls = min_logstd + softplus(ls - min_logstd)
    return mean, ls


def normalize(x, mean, std):
    """Normalizer.process, mirl/processors/normalizer.py:17-18."""
    return (x - mean) / (std + NORM_EPS)


def recover(x, mean, std):
    """Normalizer.recover, mirl/processors/normalizer.py:20-21."""
    return x * (std + NORM_EPS) + mean


# --------------------------------------------------------------------------- done functions
def done_function(kind, next_obs):
    """mirl/environments/reward_done_functions/{cheetah,hopper,walker2d,mb_ant,mb_humanoid,
    inverted_pendulum}.py ``done_function(..., 'torch')``; returns float ``[B,1]``."""
    if kind == "cheetah":                                        # cheetah.py:8-14
        return torch.zeros_like(next_obs[..., 0:1])
    if kind == "hopper":                                         # hopper.py:8-28
        height, ang = next_obs[..., 0], next_obs[..., 1]
        finite = torch.isfinite(next_obs).all(-1)
        bounded = (torch.abs(next_obs[..., 1:]) < 100).all(-1)
        live = finite & bounded & (height > 0.7) & (torch.abs(ang) < 0.2)
        return (~live)[..., None].float()
    if kind == "walker2d":                                       # walker2d.py:8-19
        height, angle = next_obs[..., 0:1], next_obs[..., 1:2]
        live = (height > 0.8) & (height < 2.0) & (angle > -1.0) & (angle < 1.0)
        return (~live).float()
    if kind == "mb_ant":                                         # mb_ant.py:8-24
        height = next_obs[..., 0]
        finite = torch.isfinite(next_obs).all(-1)
        live = finite & (height >= 0.2) & (height <= 1)
        return (~live)[..., None].float()
    if kind == "mb_humanoid":                                    # mb_humanoid.py:8-16
        q = next_obs[..., 0:1]
        return ((q < 1.0) | (q > 2.0)).float()
    if kind == "inverted_pendulum":                              # inverted_pendulum.py:8-24
        ang = next_obs[..., 1]
        finite = torch.isfinite(next_obs).all(-1)
        live = finite & (ang <= 0.2) & (ang >= -0.2)
        return (~live)[..., None].float()
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------- rollout
def ensemble_distribution(p, obs, act):
    """All-member normalised mean / logstd: the first half of PEModelModule.forward
    (pe_model.py:160-171) behind Model.step's input normalisation (base_model.py:56-63).
    Returns ``mean, logstd`` of shape ``[E,B,D]`` (D = obs+1; column D-1 is the reward)."""
    x = torch.cat([normalize(obs, p["obs_mean"], p["obs_std"]),
                   normalize(act, p["act_mean"], p["act_std"])], dim=-1)
    out = mlp_forward(p["W"], p["b"], x, p.get("act", "swish"))
    return mean_logstd(out, p["max_logstd"], p["min_logstd"])


def rollout_step(p, obs, act, idx, eps, done_kind="cheetah"):
    """PEModel.step without ``return_distribution``.

    pe_model.py:153-181 (gather the drawn member per row, ``eps*std+mean``, split) +
    base_model.py:64-76 (de-normalise delta, ``next_obs = obs + delta``, done function).
    ``idx [B]`` (int64) and ``eps [B,D]`` are what the reference draws from the global NumPy /
    torch generators (``np.random.choice(elite_indices, B)`` then ``torch.randn_like``).
    Evaluates ALL members like the reference does.  Returns next_obs, reward, done.
    """
    mean, ls = ensemble_distribution(p, obs, act)
    std = torch.exp(ls)
    ar = torch.arange(obs.shape[0])
    m, s = mean[idx, ar], std[idx, ar]
    sample = eps * s + m
    O = obs.shape[-1]
    delta = recover(sample[..., :O], p["delta_mean"], p["delta_std"])
    next_obs = obs + delta
    reward = sample[..., -1:]
    done = done_function(done_kind, next_obs)
    return next_obs, reward, done


def step_info(p, obs, act, elite_indices):
    """``info`` of PEModel.step(return_distribution=True): pe_model.py:183-188,255-282.
    Elites in the order of ``elite_indices`` (rank order)."""
    mean, ls = ensemble_distribution(p, obs, act)
    std = torch.exp(ls)
    O = obs.shape[-1]
    e = torch.as_tensor(np.asarray(elite_indices), dtype=torch.long)
    scale = p["delta_std"] + NORM_EPS
    d_mean = recover(mean[e][..., :O], p["delta_mean"], p["delta_std"])
    d_std = std[e][..., :O] * scale
    return {
        "ensemble_delta_mean": d_mean,
        "ensemble_delta_std": d_std,
        "ensemble_next_obs_mean": d_mean + obs[None, :, :],
        "ensemble_next_obs_std": d_std,
        "ensemble_reward_mean": mean[e][..., -1:],
        "ensemble_reward_std": std[e][..., -1:],
    }


def mopo_penalty(next_obs_mean, next_obs_std, kind="total_var"):
    """MOPOCollector.add_sample penalty, mirl/agents/mbpo/mopo_collector.py:42-53."""
    if kind == "max_var":
        pen = torch.norm(next_obs_std, dim=-1, keepdim=True)
        return torch.max(pen, dim=0)[0]
    if kind == "total_var":
        mean_var = torch.mean(next_obs_std ** 2, dim=0)
        var_mean = torch.var(next_obs_mean, dim=0)       # unbiased
        return torch.sum(mean_var + var_mean, dim=-1, keepdim=True)
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------- loss + grads
def _bcast(x, E):
    return x if x.dim() == 3 else x.unsqueeze(0).expand(E, *x.shape)


def model_loss(p, obs, act, delta, reward, done, ignore_terminal_state=False,
               weight_decay=(2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4), bound_reg_coef=2e-2,
               reward_coef=1.0):
    """PEModel.compute_loss -> PEModelModule.compute_loss.

    pe_model.py:285-307 (normalise obs/action/delta), :71-96 (loss assembly), :110-143
    (log_prob), :16-22 (gaussian_log_prob), mlp.py:170-178 + linear.py:161-165 (L2 term),
    pe_model.py:105-108 (bound term).  Inputs ``[E,B,.]`` or ``[B,.]``.
    Returns ``loss`` (scalar) and ``ensemble_loss [E]``.
    """
    W, b = p["W"], p["b"]
    E = W[0].shape[0]
    O = obs.shape[-1]
    x = torch.cat([normalize(obs, p["obs_mean"], p["obs_std"]),
                   normalize(act, p["act_mean"], p["act_std"])], dim=-1)
    t = normalize(delta, p["delta_mean"], p["delta_std"])
    out = mlp_forward(W, b, x, p.get("act", "swish"))
    mean, ls = mean_logstd(out, p["max_logstd"], p["min_logstd"])

    def log_prob(mu, log_std, target):
        log_var = log_std * 2
        inv_var = torch.exp(-log_var)
        return -0.5 * (LOG2PI + log_var + (target - mu) ** 2 * inv_var)

    s_lp = log_prob(mean[..., :O], ls[..., :O], t).sum(-1, keepdim=True)
    r_lp = log_prob(mean[..., -1:], ls[..., -1:], reward)
    if ignore_terminal_state:
        s_lp = s_lp * (1 - done)
        r_lp = r_lp * (1 - done)
    s_loss = -s_lp.mean([1, 2])
    r_loss = reward_coef * -r_lp.mean([1, 2])
    ensemble_loss = s_loss + r_loss
    l2 = sum(sum((W[l][e] ** 2).sum() * weight_decay[l] * 0.5 for e in range(E))
             for l in range(len(W)))
    bound = bound_reg_coef * torch.sum(p["max_logstd"].sum(0) - p["min_logstd"].sum(0))
    return ensemble_loss.sum(0) + l2 + bound, ensemble_loss


def model_loss_grads(p, obs, act, delta, reward, done, ignore_terminal_state=False,
                     weight_decay=(2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4), bound_reg_coef=2e-2,
                     reward_coef=1.0):
    """Analytic gradient of ``model_loss`` (what ``loss.backward()`` produces in
    pe_model_trainer.py:97-98), written out by hand -- SURVEY.md Appendix A.3.
    Returns dict with ``W``/``b`` lists and ``max_logstd``/``min_logstd`` grads, plus the loss.
    """
    W, b = p["W"], p["b"]
    L, E = len(W), W[0].shape[0]
    O = obs.shape[-1]
    assert p.get("act", "swish") == "swish"
    x = torch.cat([normalize(obs, p["obs_mean"], p["obs_std"]),
                   normalize(act, p["act_mean"], p["act_std"])], dim=-1)
    x = _bcast(x, E)
    t = _bcast(torch.cat([normalize(delta, p["delta_mean"], p["delta_std"]), reward], -1), E)
    B = x.shape[1]
    hs, zs = [x], []
    for l in range(L):
        z = hs[-1].matmul(W[l]) + b[l]
        zs.append(z)
        if l < L - 1:
            hs.append(z * torch.sigmoid(z))
    out = zs[-1]
    D = out.shape[-1] // 2
    mean, raw = out[..., :D], out[..., D:]
    mx, mn = p["max_logstd"], p["min_logstd"]
    ls1 = mx - softplus(mx - raw)
    ls = mn + softplus(ls1 - mn)
    coef = torch.ones(D)
    coef[-1] = reward_coef
    mask = (1 - _bcast(done, E)) if ignore_terminal_state else torch.ones(E, B, 1)
    inv_var = torch.exp(-2 * ls)
    diff = t - mean
    nll = 0.5 * (LOG2PI + 2 * ls + diff * diff * inv_var)
    ensemble_loss = (mask * (nll * coef)).sum(-1).mean(1)
    g_mean = -mask * diff * inv_var * coef / B
    g_ls = mask * (1 - diff * diff * inv_var) * coef / B
    s_a = torch.sigmoid(mx - raw)            # d softplus(mx-raw)/d(mx-raw)
    s_b = torch.sigmoid(ls1 - mn)            # d softplus(ls1-mn)/d(ls1-mn)
    s_a = torch.where(mx - raw > 20.0, torch.ones_like(s_a), s_a)
    s_b = torch.where(ls1 - mn > 20.0, torch.ones_like(s_b), s_b)
    g_raw = g_ls * s_b * s_a
    g_max = (g_ls * s_b * (1 - s_a)).sum(1, keepdim=True) + bound_reg_coef
    g_min = (g_ls * (1 - s_b)).sum(1, keepdim=True) - bound_reg_coef
    g_out = torch.cat([g_mean, g_raw], -1)
    gW, gb = [None] * L, [None] * L
    g = g_out
    for l in range(L - 1, -1, -1):
        gW[l] = hs[l].transpose(1, 2).matmul(g) + weight_decay[l] * W[l]
        gb[l] = g.sum(1, keepdim=True)
        if l > 0:
            gh = g.matmul(W[l].transpose(1, 2))
            z = zs[l - 1]
            sg = torch.sigmoid(z)
            g = gh * (sg * (1 + z * (1 - sg)))
    l2 = sum((W[l] ** 2).sum() * weight_decay[l] * 0.5 for l in range(L))
    bound = bound_reg_coef * (mx.sum() - mn.sum())
    loss = ensemble_loss.sum() + l2 + bound
    return {"W": gW, "b": gb, "max_logstd": g_max, "min_logstd": g_min,
            "loss": loss, "ensemble_loss": ensemble_loss}


def adam_update(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, beta1=0.9, beta2=0.999,
                eps=1e-8):
    """torch.optim.Adam single-tensor rule as configured at pe_model_trainer.py:47-51
    (no weight decay, no amsgrad); ``step`` is the 1-based step count.  In place."""
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / bc1)


def trainable(p):
    """Flat list of the trainable tensors of a packed parameter dict."""
    return list(p["W"]) + list(p["b"]) + [p["max_logstd"], p["min_logstd"]]


def new_adam_state(p):
    ts = trainable(p)
    return {"step": 0, "m": [torch.zeros_like(t) for t in ts],
            "v": [torch.zeros_like(t) for t in ts]}


def train_step(p, opt, batch, lr=1e-3, **loss_kwargs):
    """One PEModelTrainer.train_from_torch_batch (pe_model_trainer.py:85-106): loss,
    gradients, Adam.  Mutates ``p`` and ``opt``.  Returns (loss, ensemble_loss[E])."""
    g = model_loss_grads(p, batch["observations"], batch["actions"], batch["deltas"],
                         batch["rewards"], batch["terminals"], **loss_kwargs)
    grads = list(g["W"]) + list(g["b"]) + [g["max_logstd"], g["min_logstd"]]
    opt["step"] += 1
    for t, gr, m, v in zip(trainable(p), grads, opt["m"], opt["v"]):
        adam_update(t, gr, m, v, opt["step"], lr=lr)
    return g["loss"], g["ensemble_loss"]


def rank_elites(eval_loss, allow_to_use, elite_number):
    """PEModel.remember_loss, pe_model.py:234-237."""
    rank = np.argsort(np.asarray(eval_loss)[:allow_to_use])
    return rank, list(rank[:elite_number])
