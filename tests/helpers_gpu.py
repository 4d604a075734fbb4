"""Builders for the GPU parity tests: B200 classes loaded with the deterministic case weights."""
import numpy as np
import torch

from tests.golden import cases
from tests.util import T


class FakeEnv:
    """What the path reads from an env (base_model.py:18-27, base_value.py:5-7)."""

    class _Space:
        def __init__(self, n):
            self.shape = (n,)

    def __init__(self, env_name="cheetah", obs_dim=cases.OBS, act_dim=cases.ACT):
        self.env_name = env_name
        self.observation_space = FakeEnv._Space(obs_dim)
        self.action_space = FakeEnv._Space(act_dim)


def make_pe_model(cfg, seed, env="cheetah", device="cuda", kernel=None, obs=cases.OBS, act=cases.ACT):
    import mirl_b200
    p = cases.pe_params(seed, cfg, obs=obs, act=act)
    m = mirl_b200.B200PEModel(FakeEnv(env, obs, act), known=["done"], ensemble_size=cfg["E"],
                              elite_number=cfg["elites"], hidden_layers=list(cfg["hidden"]),
                              activation=cfg["act"], connection="simple", reward_coef=1,
                              bound_reg_coef=2e-2,
                              weight_decay=list(cases.WEIGHT_DECAY)[:len(cfg["hidden"]) + 1],
                              kernel=kernel)
    m.load_state_dict({k: T(v) for k, v in cases.pe_state_dict(p).items()})
    return m.to(device), p


def make_critic(kind, seed, device="cuda"):
    import mirl_b200
    c = cases.REDQ if kind == "redq" else cases.TQC
    W, b = cases.mlp_params(seed, c["E"], [cases.OBS + cases.ACT] + c["hidden"] + [c["out"]])
    if kind == "redq":
        q = mirl_b200.B200EnsembleQValue(FakeEnv(), ensemble_size=c["E"], hidden_layers=list(c["hidden"]),
                                         activation=c["act"])
    else:
        q = mirl_b200.B200TQCQValue(FakeEnv(), ensemble_size=c["E"], number_head=c["out"],
                                    hidden_layers=list(c["hidden"]), activation=c["act"])
    q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W, b).items()})
    return q.to(device), (W, b)


def done_guard_mismatches(kind, next_obs_ref, done_ref, done_got, band=1e-3):
    """Rows whose done flag differs AND whose reference next_obs is farther than ``band`` from
    every threshold of the done function (a flip inside the band is a legitimate fp32 rounding
    effect of a different summation order; outside it is a bug).  Returns (outside, inside)."""
    no = np.asarray(next_obs_ref, dtype=np.float64)
    diff = (np.asarray(done_ref) != np.asarray(done_got)).reshape(-1)
    thr = {"hopper": [(0, 0.7), (1, 0.2), (1, -0.2)], "walker2d": [(0, 0.8), (0, 2.0), (1, -1.0), (1, 1.0)],
           "mb_ant": [(0, 0.2), (0, 1.0)], "mb_humanoid": [(0, 1.0), (0, 2.0)],
           "inverted_pendulum": [(1, 0.2), (1, -0.2)], "cheetah": []}[kind]
    near = np.zeros(no.shape[0], dtype=bool)
    for col, t in thr:
        near |= np.abs(no[:, col] - t) <= band * max(1.0, abs(t))
    if kind == "hopper":
        near |= (np.abs(np.abs(no[:, 1:]) - 100) <= 0.1).any(1)
    return int((diff & ~near).sum()), int((diff & near).sum())
