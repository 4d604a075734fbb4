"""Deterministic synthetic inputs shared by the golden generator, the oracle tests and the GPU
parity tests.  Everything comes from ``np.random.RandomState`` (a frozen legacy stream), so a
fixture only has to store the reference's OUTPUTS; weights and inputs are rebuilt from seeds.
"""
import numpy as np

OBS, ACT = 17, 6
MBPO = dict(E=7, elites=5, hidden=[200, 200, 200, 200], act="swish")
MOPO = dict(E=10, elites=7, hidden=[200, 200, 200, 200], act="swish")
SMALL = dict(E=3, elites=2, hidden=[16, 16], act="swish")
REDQ = dict(E=10, hidden=[256, 256], act="relu", out=1)
TQC = dict(E=5, hidden=[512, 512, 512], act="relu", out=25)
WEIGHT_DECAY = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4]          # configs/mbpo.json:104


def mlp_params(seed, E, sizes, bias_scale=0.1):
    """Weights ``[E,in,out] ~ N(0,1/in)`` (keeps activations O(1) like the reference's
    orthogonal init) and small non-zero biases ``[E,1,out]``."""
    rs = np.random.RandomState(seed)
    W, b = [], []
    for i in range(len(sizes) - 1):
        W.append((rs.standard_normal((E, sizes[i], sizes[i + 1])) / np.sqrt(sizes[i]))
                 .astype(np.float32))
        b.append((bias_scale * rs.standard_normal((E, 1, sizes[i + 1]))).astype(np.float32))
    return W, b


def pe_params(seed, cfg, obs=OBS, act=ACT):
    """Packed parameter dict of a probabilistic-ensemble model (numpy, fp32)."""
    E, D = cfg["E"], obs + 1
    W, b = mlp_params(seed, E, [obs + act] + list(cfg["hidden"]) + [2 * D])
    rs = np.random.RandomState(seed + 1)
    # log-std head biased low so both soft clamps are exercised (raw in about [-8, 2])
    b[-1][:, :, D:] += -3.0
    W[-1][:, :, D:] *= 3.0
    return {
        "W": W, "b": b, "act": cfg["act"],
        "max_logstd": (0.25 + 0.2 * rs.standard_normal((E, 1, D))).astype(np.float32),
        "min_logstd": (-5.0 + 0.5 * rs.standard_normal((E, 1, D))).astype(np.float32),
        "obs_mean": rs.standard_normal(obs).astype(np.float32),
        "obs_std": rs.uniform(0.5, 1.5, obs).astype(np.float32),
        "act_mean": (0.1 * rs.standard_normal(act)).astype(np.float32),
        "act_std": rs.uniform(0.5, 1.5, act).astype(np.float32),
        "delta_mean": (0.1 * rs.standard_normal(obs)).astype(np.float32),
        "delta_std": rs.uniform(0.05, 0.5, obs).astype(np.float32),
    }


def rollout_inputs(seed, B, obs=OBS, act=ACT, env="cheetah"):
    rs = np.random.RandomState(seed)
    o = rs.standard_normal((B, obs)).astype(np.float32)
    if env in ("hopper", "inverted_pendulum"):
        o[:, 0] = rs.uniform(0.6, 1.4, B)      # height around the 0.7 threshold
        o[:, 1] = rs.uniform(-0.25, 0.25, B)   # angle around the +-0.2 threshold
    elif env == "walker2d":
        o[:, 0] = rs.uniform(0.7, 2.1, B)
        o[:, 1] = rs.uniform(-1.1, 1.1, B)
    elif env == "mb_ant":
        o[:, 0] = rs.uniform(0.1, 1.1, B)
    elif env == "mb_humanoid":
        o[:, 0] = rs.uniform(0.9, 2.1, B)
    a = rs.uniform(-1, 1, (B, act)).astype(np.float32)
    return o, a


def train_batch(seed, E, B, obs=OBS, act=ACT, shared=False):
    """Per-member bootstrap batch ``[E,B,.]`` (pe_model_trainer.py:268-270) or a shared eval
    batch ``[B,.]`` (:108-115)."""
    rs = np.random.RandomState(seed)
    lead = (B,) if shared else (E, B)
    return {
        "observations": rs.standard_normal(lead + (obs,)).astype(np.float32),
        "actions": rs.uniform(-1, 1, lead + (act,)).astype(np.float32),
        "deltas": (0.2 * rs.standard_normal(lead + (obs,))).astype(np.float32),
        "rewards": rs.standard_normal(lead + (1,)).astype(np.float32),
        "terminals": (rs.uniform(size=lead + (1,)) < 0.1).astype(np.float32),
    }


def critic_inputs(seed, B, obs=OBS, act=ACT):
    rs = np.random.RandomState(seed)
    return {
        "obs": rs.standard_normal((B, obs)).astype(np.float32),
        "act": rs.uniform(-1, 1, (B, act)).astype(np.float32),
        "rewards": rs.standard_normal((B, 1)).astype(np.float32),
        "terminals": (rs.uniform(size=(B, 1)) < 0.1).astype(np.float32),
        "log_prob": rs.standard_normal((B, 1)).astype(np.float32),
    }


def pe_state_dict(p, module_prefix="module."):
    """Reference ``PEModel.state_dict()`` key layout for packed params."""
    sd = {
        "obs_processor.mean": p["obs_mean"], "obs_processor.std": p["obs_std"],
        "action_processor.mean": p["act_mean"], "action_processor.std": p["act_std"],
        "delta_processor.mean": p["delta_mean"], "delta_processor.std": p["delta_std"],
    }
    E = p["W"][0].shape[0]
    for e in range(E):
        sd["%smaxlogstd_net%d" % (module_prefix, e)] = p["max_logstd"][e]
        sd["%sminlogstd_net%d" % (module_prefix, e)] = p["min_logstd"][e]
    sd.update(mlp_state_dict(p["W"], p["b"], module_prefix))
    return sd


def mlp_state_dict(W, b, module_prefix="module."):
    sd = {}
    for l in range(len(W)):
        for e in range(W[l].shape[0]):
            sd["%snet.%d.weight_net%d" % (module_prefix, 2 * l, e)] = W[l][e]
            sd["%snet.%d.bias_net%d" % (module_prefix, 2 * l, e)] = b[l][e]
    return sd


# ---- SimplePool fixture (tests/golden/pool.npz) ------------------------------------------------
POOL_ADDS = [300, 450, 200, 400, 1, 650]        # capacity 1000: the 3rd add reaches the end, later ones wrap


def pool_samples(seed, n, obs=OBS, act=ACT):
    rs = np.random.RandomState(seed)
    return {
        "observations": rs.standard_normal((n, obs)).astype(np.float32),
        "actions": rs.uniform(-1, 1, (n, act)).astype(np.float32),
        "next_observations": rs.standard_normal((n, obs)).astype(np.float32),
        "rewards": rs.standard_normal((n, 1)).astype(np.float32),
        "terminals": (rs.uniform(size=(n, 1)) < 0.1).astype(np.float32),
    }


# ---- learnable synthetic dynamics (train-loop / ExtraFieldPool / collector fixtures) -------------
def dyn_samples(seed, n, obs=OBS, act=ACT, env="cheetah"):
    """Transitions of a fixed smooth synthetic system (so that model training has something to
    learn and the early-stopping counters move): delta = 0.1 tanh(o A + a B) + noise."""
    rs = np.random.RandomState(seed)
    ms = np.random.RandomState(4242)
    Aw = (0.6 * ms.standard_normal((obs, obs)) / np.sqrt(obs)).astype(np.float32)
    Bw = (0.5 * ms.standard_normal((act, obs))).astype(np.float32)
    rw = (ms.standard_normal(obs + act) / 3.0).astype(np.float32)
    o, a = rollout_inputs(seed + 1, n, obs, act, env)
    delta = (0.1 * np.tanh(o @ Aw + a @ Bw) + 0.02 * rs.standard_normal((n, obs))).astype(np.float32)
    r = (np.tanh(np.concatenate([o, a], 1) @ rw) + 0.05 * rs.standard_normal(n)).astype(np.float32)
    return {
        "observations": o, "actions": a, "next_observations": (o + delta).astype(np.float32),
        "rewards": r[:, None].astype(np.float32),
        "terminals": (rs.uniform(size=(n, 1)) < 0.05).astype(np.float32),
    }


# PEModelTrainer.train_with_pool fixture (tests/golden/pe_trainloop_*.npz): two pool fills, two calls
TRAINLOOP = {
    "mbpo": dict(cfg=MBPO, seed=11, adds=[(7001, 1200), (7002, 500)], batch_size=256, report_freq=8,
                 max_not_improve=2, init_model_train_step=48, max_step_per_train=40, max_valid=300, lr=1e-3,
                 rng_seed=80),
    # a large step size makes the hold-out loss non-monotonic: non-improvements, per-member early
    # stopping and the restore of an EARLIER best snapshot are all exercised
    "small": dict(cfg=SMALL, seed=51, adds=[(7101, 300), (7102, 200)], batch_size=32, report_freq=10,
                  max_not_improve=2, init_model_train_step=300, max_step_per_train=60, max_valid=100,
                  lr=0.1, rng_seed=76),
}


class TanhPolicy:
    """Stand-in for ``GaussianPolicy.action`` in the collector fixtures: fixed weights,
    ``a = tanh(obs M + 0.3 eps)`` with ``eps [B,A]`` drawn from the global torch generator -- the
    same place in the RNG consumption order as the reference policy's draw (mbpo_collector.py:48,
    gaussian.py:199-204).  ``noise_source(B, A)`` replays recorded draws instead."""

    def __init__(self, seed, obs=OBS, act=ACT, noise_source=None):
        import torch
        self.M = torch.from_numpy((np.random.RandomState(seed).standard_normal((obs, act)) /
                                   np.sqrt(obs)).astype(np.float32))
        self.act = act
        self.noise_source = noise_source
        self.drawn = []

    def action(self, obs, **kwargs):
        import torch
        B = obs.shape[0]
        if self.noise_source is not None:
            eps = self.noise_source(B, self.act).to(obs.device)
        else:
            eps = torch.randn(B, self.act, device=obs.device)
            self.drawn.append(eps.cpu().numpy().copy())
        return torch.tanh(obs @ self.M.to(obs.device) + 0.3 * eps), {}


# MBPOCollector / MOPOCollector fixtures (tests/golden/collector_*.npz)
COLLECT = {
    "mbpo": dict(cfg=MBPO, mseed=11, env="hopper", pool_seed=8001, pool_n=1500, number_sample=600,
                 batch_size=400, depth=5, save_n=2, loss=[0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4]),
    "mopo": dict(cfg=MOPO, mseed=31, env="hopper", pool_seed=8002, pool_n=800, number_sample=300,
                 batch_size=200, depth=3, save_n=2, penalty_coef=0.7, penalty_type="total_var",
                 loss=[0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4, 0.05, 0.8, 0.6]),
}


def critic_q_target(x):
    """Synthetic TD target ``[B,1]`` for the critic-update fixture (the real one comes from
    SACAgent.compute_q_target, which has its own fixture)."""
    return (3.0 * x["rewards"] + x["log_prob"]).astype(np.float32)


def tqc_atoms_target(x, seed, n_target=115):
    """Synthetic target atoms ``[B, n_target]`` for the TQC critic-update fixture (ascending per row, like
    TQCAgent.compute_q_target's sorted-and-truncated atoms; spread wide enough that both Huber branches occur)."""
    rs = np.random.RandomState(seed)
    B = x["rewards"].shape[0]
    spread = np.sort(rs.standard_normal((B, n_target)), axis=1)
    return (x["rewards"] + 0.5 * x["log_prob"] + 1.5 * spread).astype(np.float32)


def sac_batch(seed, B, obs=OBS, act=ACT):
    """A replay batch with the keys SACAgent.train_from_torch_batch reads (sac_agent.py:153-157)."""
    rs = np.random.RandomState(seed)
    return {
        "observations": rs.standard_normal((B, obs)).astype(np.float32),
        "actions": rs.uniform(-1, 1, (B, act)).astype(np.float32),
        "rewards": rs.standard_normal((B, 1)).astype(np.float32),
        "terminals": (rs.uniform(size=(B, 1)) < 0.1).astype(np.float32),
        "next_observations": rs.standard_normal((B, obs)).astype(np.float32),
    }


def actor_noise(step, B=256, act=ACT):
    """The standard-normal draw ``policy.action`` consumed in gen_golden.gen_actor_update: ``dist.rsample()`` after
    ``torch.manual_seed(980 + step)`` is ``mean + std * torch.randn(B, A)`` of the same generator state."""
    import torch
    torch.manual_seed(980 + step)
    return torch.randn(B, act).numpy()


def policy_state_dict(W, b):
    """Reference ``GaussianPolicy.state_dict()`` keys (a plain MLP: ``module.net.{2l}.weight`` / ``.bias``)."""
    sd = {}
    for l in range(len(W)):
        sd["module.net.%d.weight" % (2 * l)] = W[l][0]
        sd["module.net.%d.bias" % (2 * l)] = b[l][0]
    return sd
