"""CPU oracle for the two loops around the ensemble kernels.  TEST INFRASTRUCTURE ONLY.

Restatements (NumPy / torch CPU, over ``oracle/pe_oracle.py``) of
  * ``PEModelTrainer.train_with_pool`` / ``sufficiently_train`` / ``split_dataset`` / ``eval_with_dataset``
    (mirl/agents/mbpo/pe_model_trainer.py:108-121,138-182,212-293) with the ``ExtraFieldPool`` protocol it relies
    on (mirl/pools/extra_field_pool.py:47-66, flag_pool.py:33-54, utils/mean_std.py:12-33) and
    ``PEModel.remember_loss`` (pe_model.py:234-237);
  * ``MBPOCollector.imagine`` / ``add_sample`` (mirl/agents/mbpo/mbpo_collector.py:34-97) and the MOPO reward
    penalty (mopo_collector.py:42-53).
Pinned by ``tests/golden/pe_trainloop_*.npz`` and ``tests/golden/collector_*.npz`` (outputs of the unmodified
reference, ``tests/golden/gen_golden.py``).  Same import rules as the other oracle modules.
"""
import numpy as np
import torch

from . import pe_oracle as po

FIELDS = ("observations", "actions", "deltas", "rewards", "terminals")


class RunningMeanStd(object):
    """mirl/utils/mean_std.py:3-33 (float64 state; batch moments with NumPy's float32 reductions)."""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.std = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        bm, bv, n = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
        delta = bm - self.mean
        tot = self.count + n
        m2 = self.var * self.count + bv * n + np.square(delta) * self.count * n / tot
        self.mean, self.var, self.count = self.mean + delta * n / tot, m2 / tot, tot
        self.std = np.sqrt(self.var)


class ExtraFieldPoolModel(object):
    """The part of ExtraFieldPool train_with_pool uses, for a pool that never wraps: ordered storage, running
    statistics per field, the 'deltas' extra field and the unprocessed-rows flag."""

    def __init__(self):
        self.data = {}
        self.stats = {}
        self.unprocessed = 0

    def _append(self, k, v):
        v = np.asarray(v, dtype=np.float32)
        self.data[k] = v if k not in self.data else np.concatenate([self.data[k], v])
        self.stats.setdefault(k, RunningMeanStd(v.shape[1:])).update(v)

    def add_samples(self, samples):
        n = len(samples["observations"])
        for k in ("observations", "next_observations", "actions", "rewards", "terminals"):   # simple_pool.py:57-81
            self._append(k, samples[k])
        self.unprocessed += n

    def compute_deltas(self):
        """pe_model_trainer.py:140-143."""
        n = self.unprocessed
        size = len(self.data["observations"])
        d = self.data["next_observations"][size - n:] - self.data["observations"][size - n:]
        self._append("deltas", d)
        self.unprocessed = 0

    def mean_std(self):
        return {k: (s.mean, s.std) for k, s in self.stats.items()}


def set_mean_std(p, ms):
    """PEModelTrainer.set_mean_std (pe_model_trainer.py:123-136) -> Normalizer.set_mean_std_np."""
    for field, name in (("observations", "obs"), ("actions", "act"), ("deltas", "delta")):
        p[name + "_mean"] = torch.from_numpy(np.asarray(ms[field][0])).float()
        p[name + "_std"] = torch.from_numpy(np.asarray(ms[field][1])).float()


def eval_with_dataset(p, ds, **loss_kwargs):
    """pe_model_trainer.py:108-121: hold-out loss of every member (compute_loss default ignore_terminal_state=True)."""
    _, el = po.model_loss(p, ds["observations"], ds["actions"], ds["deltas"], ds["rewards"], ds["terminals"],
                          ignore_terminal_state=True, **loss_kwargs)
    return [float(v) for v in el.numpy()]


def member_tensors(p, e):
    return [t[e] for t in po.trainable(p)]


def train_with_pool(p, opt, pool, max_step, cfg, elite_number, lr=1e-3, record=None, **loss_kwargs):
    """PEModelTrainer.train_with_pool -> sufficiently_train (pe_model_trainer.py:138-170,212-293).  ``cfg`` holds
    batch_size / report_freq / max_not_improve / max_valid / valid_ratio.  Consumes the global NumPy generator
    like the reference (permutation, then the bootstrap table).  Returns a dict of results."""
    pool.compute_deltas()
    set_mean_std(p, pool.mean_std())
    E = p["W"][0].shape[0]
    n_sample = len(pool.data["observations"])
    valid_length = int(np.ceil(min(cfg["max_valid"], cfg.get("valid_ratio", 0.2) * n_sample)))   # pools/utils.py:60-79
    train_length = n_sample - valid_length
    indices = np.random.permutation(n_sample)
    tr_i, va_i = indices[:train_length], indices[-valid_length:]
    train = {k: pool.data[k][tr_i] for k in FIELDS}
    evald = {k: torch.from_numpy(pool.data[k][va_i]) for k in FIELDS}
    ensemble_index = np.random.randint(train_length, size=(E, train_length))                    # :179
    bs, rf, mni = cfg["batch_size"], cfg["report_freq"], cfg["max_not_improve"]
    total, best = 0, [None] * E
    min_loss, not_improve = None, None
    go = True
    evals, trains = [], []
    while True:
        for j in range(0, train_length, bs):
            if total % rf == 0:
                el = eval_with_dataset(p, evald, **loss_kwargs)
                evals.append(el)
                if total == 0:
                    min_loss, not_improve = list(el), [0] * E
                    for i in range(E):
                        best[i] = [t.clone() for t in member_tensors(p, i)]
                else:
                    for i, l in enumerate(el):
                        if l < min_loss[i]:
                            min_loss[i], not_improve[i] = l, 0
                            best[i] = [t.clone() for t in member_tensors(p, i)]
                        else:
                            not_improve[i] += 1
                go = any(mni < 0 or n < mni for n in not_improve)
            if total >= max_step:
                go = False
            if not go:
                break
            ind = ensemble_index[:, j:j + bs]
            batch = {k: torch.from_numpy(train[k][ind]) for k in FIELDS}
            loss, el_t = po.train_step(p, opt, batch, lr=lr, **loss_kwargs)
            trains.append([float(v) for v in el_t.numpy()] + [float(loss)])
            total += 1
        if not go:
            break
    for i in range(E):                                             # load_best (:282-284)
        for t, s in zip(member_tensors(p, i), best[i]):
            t.copy_(s)
    final = eval_with_dataset(p, evald, **loss_kwargs)
    evals.append(final)
    rank, elites = po.rank_elites(final, E, elite_number)
    return {"eval": np.array(evals), "train": np.array(trains), "train_step": total, "min_loss": final,
            "rank": rank, "elite_indices": elites, "lengths": (train_length, valid_length),
            "ensemble_index": ensemble_index}


def imagine(p, elite_indices, start_states, policy_action, noise_source, env, depth, batch_size,
            penalty=None):
    """MBPOCollector.imagine (mbpo_collector.py:34-54) + add_sample (:69-88).  ``start_states`` is the
    ``random_batch_torch`` draw; ``policy_action(o)`` -> actions; ``noise_source(B)`` -> the torch.randn draw the
    model consumes; member indices come from the global NumPy generator like pe_model.py:169.
    ``penalty=(kind, coef)`` applies MOPOCollector's reward penalty (mopo_collector.py:42-53).
    Returns the pool contents in insertion order."""
    out = {k: [] for k in ("observations", "actions", "next_observations", "rewards", "terminals")}
    n = start_states.shape[0]
    for i in range(0, n, batch_size):
        o = start_states[i:i + min(batch_size, n - i)]
        stop = torch.zeros(o.shape[0], 1)
        for _ in range(depth):
            a = policy_action(o)
            idx = torch.from_numpy(np.random.choice(elite_indices, o.shape[0]))
            eps = noise_source(o.shape[0])
            next_o, r, d = po.rollout_step(p, o, a, idx, eps, env)
            if penalty is not None:
                info = po.step_info(p, o, a, elite_indices)
                pen = po.mopo_penalty(info["ensemble_next_obs_mean"], info["ensemble_next_obs_std"], penalty[0])
                r = r - penalty[1] * pen
            d = stop + d - stop * d
            live = (stop < 0.5).flatten()
            for k, v in (("observations", o), ("actions", a), ("next_observations", next_o), ("rewards", r),
                         ("terminals", d)):
                out[k].append(v[live].numpy())
            keep = (d < 0.5).flatten()
            o, stop = next_o[keep], d[keep]
            if len(stop) == 0 or torch.sum(stop) == len(stop):
                break
    return {k: np.concatenate(v) for k, v in out.items()}
