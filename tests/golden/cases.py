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
