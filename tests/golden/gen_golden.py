"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Container-only: needs ``/root/reference`` (imported through ``oracle/ref_shim.py``; reference
sources are never copied).  Run:  ``python -m tests.golden.gen_golden``  from the repo root.
The fixtures store only the reference's outputs (+ the NumPy/torch RNG draws it consumed);
weights and inputs are rebuilt from seeds by ``tests/golden/cases.py``.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

from oracle import ref_shim
from tests.golden import cases

HERE = os.path.dirname(os.path.abspath(__file__))
LOSS_VEC7 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4]
LOSS_VEC10 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4, 0.05, 0.8, 0.6]


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def proj(t, seed):
    """Checksum of a tensor: norm and a projection on a fixed random direction (fp64)."""
    a = t.detach().double().numpy()
    r = np.random.RandomState(seed).standard_normal(a.shape)
    return np.array([np.sqrt((a * a).sum()), (a * r).sum()])


def build_pe(cfg, seed, env="cheetah"):
    from mirl.agents.mbpo.pe_model import PEModel
    p = cases.pe_params(seed, cfg)
    m = PEModel(ref_shim.FakeEnv(env), known=["done"], ensemble_size=cfg["E"],
                elite_number=cfg["elites"], hidden_layers=list(cfg["hidden"]),
                activation=cfg["act"], connection="simple", reward_coef=1, bound_reg_coef=2e-2,
                weight_decay=list(cases.WEIGHT_DECAY)[:len(cfg["hidden"]) + 1])
    sd = {k: T(v) for k, v in cases.pe_state_dict(p).items()}
    m.load_state_dict(sd)
    return m, p


def gen_forward():
    m, _ = build_pe(cases.MBPO, 11)
    o, a = cases.rollout_inputs(12, 200)
    with torch.no_grad():
        def run(model, o, a):
            x = torch.cat([model.obs_processor.process(o), model.action_processor.process(a)], -1)
            from mirl.torch_modules.mlp import MLP
            out = MLP.forward(model.module, x)
            return model.module._get_mean_logstd(out)
        mean, ls = run(m, T(o), T(a))
        m64 = m.double()
        mean64, ls64 = run(m64, T(o).double(), T(a).double())
    np.savez(os.path.join(HERE, "pe_forward_mbpo.npz"), mean=mean.numpy(), logstd=ls.numpy(),
             mean64=mean64.numpy(), logstd64=ls64.numpy())


def gen_step():
    for env, seed in (("cheetah", 21), ("hopper", 22), ("walker2d", 23), ("mb_ant", 24),
                      ("mb_humanoid", 25), ("inverted_pendulum", 26)):
        m, _ = build_pe(cases.MBPO, 11, env)
        m.remember_loss(LOSS_VEC7)
        B = 1000
        o, a = cases.rollout_inputs(seed, B, env=env)
        np.random.seed(seed); torch.manual_seed(seed)
        with torch.no_grad():
            no, r, d, _ = m.step(T(o), T(a))
        # replay of what step() consumed from the two global generators (pe_model.py:169,176)
        np.random.seed(seed); torch.manual_seed(seed)
        idx = np.random.choice(m.elite_indices, B)
        eps = torch.randn(B, cases.OBS + 1)
        np.savez(os.path.join(HERE, "pe_step_%s.npz" % env), idx=idx.astype(np.int64),
                 eps=eps.numpy(), next_obs=no.numpy(), reward=r.numpy(), done=d.numpy(),
                 elite_indices=np.asarray(m.elite_indices, dtype=np.int64))
    # elite_indices=None before the first training (pe_model.py:167-168)
    m, _ = build_pe(cases.MBPO, 11)
    o, a = cases.rollout_inputs(27, 64)
    np.random.seed(27); torch.manual_seed(27)
    with torch.no_grad():
        no, r, d, _ = m.step(T(o), T(a))
    np.random.seed(27); torch.manual_seed(27)
    idx = np.random.choice(list(np.arange(7)), 64)
    eps = torch.randn(64, cases.OBS + 1)
    np.savez(os.path.join(HERE, "pe_step_noelite.npz"), idx=idx.astype(np.int64), eps=eps.numpy(),
             next_obs=no.numpy(), reward=r.numpy(), done=d.numpy())


def gen_dist():
    from mirl.agents.mbpo.mopo_collector import MOPOCollector
    m, _ = build_pe(cases.MOPO, 31, "hopper")
    m.remember_loss(LOSS_VEC10)
    B = 128
    o, a = cases.rollout_inputs(32, B, env="hopper")
    np.random.seed(32); torch.manual_seed(32)
    with torch.no_grad():
        no, r, d, info = m.step(T(o), T(a), return_distribution=True)
    np.random.seed(32); torch.manual_seed(32)
    idx = np.random.choice(m.elite_indices, B)
    eps = torch.randn(B, cases.OBS + 1)
    out = {k: v.numpy() for k, v in info.items()}
    means, stds = info["ensemble_next_obs_mean"], info["ensemble_next_obs_std"]
    # MOPOCollector.add_sample penalty arithmetic (mopo_collector.py:42-53), via the class
    class _Pool:
        def add_samples(self, s): self.s = s
    for kind in ("total_var", "max_var"):
        pool = _Pool()
        c = MOPOCollector(m, None, None, pool, penalty_coef=1.0, penalty_type=kind)
        from mirl.utils.logger import logger
        logger.set_log_level(logger.ERROR) if hasattr(logger, "set_log_level") else None
        stop = torch.zeros(B, 1)
        with torch.no_grad():
            try:
                c.add_sample(stop, T(o), T(a), no, r, d, info, 0, 0)
                out["penalized_reward_" + kind] = pool.s["rewards"]
            except Exception as e:            # logger side effects; fall back to the formulae
                print("add_sample failed (%s); using formulae" % e)
                if kind == "max_var":
                    pen = torch.max(torch.norm(stds, dim=-1, keepdim=True), dim=0)[0]
                else:
                    pen = torch.sum(torch.mean(stds ** 2, 0) + torch.var(means, 0), -1, keepdim=True)
                out["penalized_reward_" + kind] = (r - pen).numpy()
    np.savez(os.path.join(HERE, "pe_step_dist_mopo.npz"), idx=idx.astype(np.int64),
             eps=eps.numpy(), next_obs=no.numpy(), reward=r.numpy(), done=d.numpy(),
             elite_indices=np.asarray(m.elite_indices, dtype=np.int64), **out)


def named_trainables(m):
    return [(n, p) for n, p in m.named_parameters() if p.requires_grad]


def gen_loss():
    res = {}
    m, _ = build_pe(cases.MBPO, 11)
    names = [n for n, _ in named_trainables(m)]
    for tag, ignore in (("keep", False), ("ignore", True)):
        bt = {k: T(v) for k, v in cases.train_batch(41, 7, 256).items()}
        m.zero_grad()
        loss, el = m.compute_loss(bt["observations"], bt["actions"], bt["deltas"], bt["rewards"],
                                  bt["terminals"], ignore)
        loss.backward()
        res["loss_" + tag] = loss.detach().numpy()
        res["ensemble_loss_" + tag] = el.detach().numpy()
        res["grad_proj_" + tag] = np.stack([proj(p.grad, 1000 + i)
                                            for i, (_, p) in enumerate(named_trainables(m))])
    bs = {k: T(v) for k, v in cases.train_batch(42, 7, 500, shared=True).items()}
    with torch.no_grad():
        loss, el = m.compute_loss(bs["observations"], bs["actions"], bs["deltas"], bs["rewards"],
                                  bs["terminals"])          # eval default ignore_terminal_state=True
    res["loss_shared"] = loss.numpy(); res["ensemble_loss_shared"] = el.numpy()
    np.savez(os.path.join(HERE, "pe_loss_mbpo.npz"), names=np.array(names), **res)

    # small config: full gradients
    m, _ = build_pe(cases.SMALL, 51)
    bt = {k: T(v) for k, v in cases.train_batch(52, 3, 64).items()}
    m.zero_grad()
    loss, el = m.compute_loss(bt["observations"], bt["actions"], bt["deltas"], bt["rewards"],
                              bt["terminals"], False)
    loss.backward()
    out = {"loss": loss.detach().numpy(), "ensemble_loss": el.detach().numpy()}
    for n, p in named_trainables(m):
        out["grad/" + n] = p.grad.numpy()
    np.savez(os.path.join(HERE, "pe_loss_small.npz"), **out)


def gen_train():
    from mirl.agents.mbpo.pe_model_trainer import PEModelTrainer
    from mirl.utils.logger import logger
    tmp = tempfile.mkdtemp()
    logger.set_snapshot_dir(tmp)
    only100 = os.environ.get("GOLDEN_TRAIN_ONLY100")
    for tag, cfg, seed, B, steps in (("mbpo", cases.MBPO, 11, 256, 10), ("small", cases.SMALL, 51, 64, 10),
                                     ("mbpo100", cases.MBPO, 11, 256, 100)):
        if only100 and steps != 100:
            continue
        m, _ = build_pe(cfg, seed)
        tr = PEModelTrainer(ref_shim.FakeEnv(), m, lr=1e-3, tb_log_freq=-1, silent=True,
                            ignore_terminal_state=False)
        out = {"loss": [], "ensemble_loss": []}
        for s in range(steps):
            bt = {k: T(v) for k, v in cases.train_batch(600 + s, cfg["E"], B).items()}
            diag = tr.train_from_torch_batch(bt)
            out["loss"].append(diag["train_loss"])
            out["ensemble_loss"].append([diag["mt/net_%d/train_loss" % i] for i in range(cfg["E"])])
            if s + 1 in (1, steps):
                nt = named_trainables(m)
                out["param_proj_%d" % (s + 1)] = np.stack([proj(p, 2000 + i) for i, (_, p) in enumerate(nt)])
                st = tr.optimizer.state
                out["m_proj_%d" % (s + 1)] = np.stack([proj(st[p]["exp_avg"], 3000 + i) for i, (_, p) in enumerate(nt)])
                out["v_proj_%d" % (s + 1)] = np.stack([proj(st[p]["exp_avg_sq"], 4000 + i) for i, (_, p) in enumerate(nt)])
                if tag == "small":
                    for n, p in nt:  # full tensors only for the small config
                        out["param_%d/%s" % (s + 1, n)] = p.detach().numpy().copy()
        out["loss"] = np.array(out["loss"], dtype=np.float64)
        out["ensemble_loss"] = np.array(out["ensemble_loss"], dtype=np.float64)
        out["names"] = np.array([n for n, _ in named_trainables(m)])
        np.savez(os.path.join(HERE, "pe_train_%s.npz" % tag), **out)


def gen_critics():
    from mirl.values.ensemble_value import EnsembleQValue, sample_from_ensemble
    from mirl.values.distributional_value import TQCQValue
    env = ref_shim.FakeEnv()
    c = cases.REDQ
    q = EnsembleQValue(env, ensemble_size=c["E"], hidden_layers=list(c["hidden"]), activation=c["act"])
    W, b = cases.mlp_params(71, c["E"], [23] + c["hidden"] + [c["out"]])
    q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W, b).items()})
    x = cases.critic_inputs(72, 256)
    o, a = T(x["obs"]), T(x["act"])
    out = {}
    with torch.no_grad():
        np.random.seed(73)
        out["bw_min"] = q.value(o, a, sample_number=2, batchwise_sample=True, mode="min")[0].numpy()
        np.random.seed(73)
        out["bw_idx"] = np.random.choice(c["E"], 2, replace=False)
        np.random.seed(74)
        out["row_min"] = q.value(o, a, sample_number=2, batchwise_sample=False, mode="min")[0].numpy()
        np.random.seed(74)
        idx = [np.random.choice(c["E"] - i, (256,)) for i in range(2)]
        out["row_raw_draws"] = np.stack(idx)
        np.random.seed(75)
        out["row3_max"] = q.value(o, a, sample_number=3, batchwise_sample=False, mode="max")[0].numpy()
        out["all_mean"] = q.value(o, a, sample_number=None, mode="mean")[0].numpy()
        out["all_min"] = q.value(o, a, sample_number=10, mode="min")[0].numpy()
        out["ensemble"] = q.value(o, a, only_return_ensemble=True)[1]["ensemble_value"].numpy()
        # SACAgent.compute_q_target arithmetic (sac_agent.py:80-85) on the batch-wise value
        alpha, disc = 0.2, 0.99
        out["td_target"] = (T(x["rewards"]) + (1. - T(x["terminals"])) * disc *
                            (T(out["bw_min"]) - alpha * T(x["log_prob"]))).numpy()
    np.savez(os.path.join(HERE, "redq.npz"), **out)

    c = cases.TQC
    q = TQCQValue(env, ensemble_size=c["E"], number_head=c["out"], hidden_layers=list(c["hidden"]),
                  activation=c["act"])
    W, b = cases.mlp_params(81, c["E"], [23] + c["hidden"] + [c["out"]])
    q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W, b).items()})
    x = cases.critic_inputs(82, 256)
    o, a = T(x["obs"]), T(x["act"])
    out = {}
    with torch.no_grad():
        v, info = q.value(o, a, percentage_drop=8, return_atoms=True)
        out["value_drop8pct"], out["atoms_drop8pct"] = v.numpy(), info["value_atoms"].numpy()
        v, info = q.value(o, a, number_drop=0, return_atoms=True, return_ensemble=True)
        out["value_drop0"], out["atoms_drop0"] = v.numpy(), info["value_atoms"].numpy()
        out["ensemble"] = q.value(o, a, only_return_ensemble=True)[1]["ensemble_atoms"].numpy()
        alpha, disc = 0.2, 0.99
        out["atoms_target"] = (T(x["rewards"]) + (1. - T(x["terminals"])) * disc *
                               (T(out["atoms_drop8pct"]) - alpha * T(x["log_prob"]))).numpy()
    np.savez(os.path.join(HERE, "tqc.npz"), **out)


def gen_keys():
    from mirl.values.ensemble_value import EnsembleQValue
    from mirl.values.distributional_value import TQCQValue
    m, _ = build_pe(cases.MBPO, 11)
    env = ref_shim.FakeEnv()
    q = EnsembleQValue(env, ensemble_size=10, hidden_layers=[256, 256], activation="relu")
    t = TQCQValue(env, ensemble_size=5, number_head=25, hidden_layers=[512] * 3, activation="relu")
    out = {n: {k: list(v.shape) for k, v in mod.state_dict().items()}
           for n, mod in (("PEModel", m), ("EnsembleQValue", q), ("TQCQValue", t))}
    out["PEModel.parameters"] = [n for n, _ in m.named_parameters()]
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(out, f, indent=0)


def gen_pool():
    """SimplePool (mirl/pools/simple_pool.py) as a ring buffer: adds that wrap, sampled batches
    (global NumPy generator), get_data ordering after the wrap, resize."""
    from mirl.pools.simple_pool import SimplePool
    env = ref_shim.FakeEnv("cheetah")
    pool = SimplePool(env, max_size=1000)
    out = {}
    np.random.seed(91)
    for i, n in enumerate(cases.POOL_ADDS):
        pool.add_samples(cases.pool_samples(910 + i, n))
        b = pool.random_batch(64)
        out["size_%d" % i] = np.array([pool.get_size(), pool._stop])
        for k, v in b.items():
            out["batch_%d_%s" % (i, k)] = v
    for k, v in pool.get_data().items():                 # order-sensitive checksums (norm, random projection)
        out["data_" + k] = proj(T(v), 7)
    pool.resize(1500)
    pool.add_samples(cases.pool_samples(990, 300))
    out["size_resized"] = np.array([pool.get_size(), pool._stop])
    for k, v in pool.get_data().items():
        out["data_resized_" + k] = proj(T(v), 8)
    b = pool.random_batch(32, keys=["observations", "rewards"])
    for k, v in b.items():
        out["batch_resized_" + k] = v
    np.savez(os.path.join(HERE, "pool.npz"), **out)


def gen_trainloop():
    """PEModelTrainer.train_with_pool (pe_model_trainer.py:138-293) on a synthetic ExtraFieldPool:
    two fills, two calls (init_model_train_step, then max_step_per_train).  Records the eval-loss
    trajectory at every report point, the not_improve counters, the train-loss curve, the step
    counts, the normaliser statistics the pool produced and the final ranking / parameters."""
    from mirl.agents.mbpo.pe_model_trainer import PEModelTrainer
    from mirl.pools.extra_field_pool import ExtraFieldPool
    from mirl.utils.logger import logger
    tmp = tempfile.mkdtemp()
    logger.set_snapshot_dir(tmp)
    for tag, c in cases.TRAINLOOP.items():
        cfg = c["cfg"]
        m, _ = build_pe(cfg, c["seed"])
        env = ref_shim.FakeEnv()
        pool = ExtraFieldPool(env, max_size=5000, compute_mean_std=True)
        pool.add_extra_fields({"deltas": {"shape": (cases.OBS,), "type": np.float32}})
        tr = PEModelTrainer(env, m, lr=c["lr"], tb_log_freq=-1, silent=True, ignore_terminal_state=False,
                            batch_size=c["batch_size"], report_freq=c["report_freq"],
                            max_not_improve=c["max_not_improve"], max_valid=c["max_valid"],
                            init_model_train_step=c["init_model_train_step"],
                            max_step_per_train=c["max_step_per_train"])
        rec = {"eval": [], "train": []}
        ev0, tr0 = tr.eval_with_dataset, tr.train_from_torch_batch

        def ev(ds, _f=ev0):
            st, el = _f(ds)
            rec["eval"].append(np.array(el, dtype=np.float64))
            return st, el

        def trn(bt, _f=tr0):
            d = _f(bt)
            rec["train"].append(np.array([d["mt/net_%d/train_loss" % i] for i in range(cfg["E"])] +
                                         [d["train_loss"]], dtype=np.float64))
            return d
        tr.eval_with_dataset, tr.train_from_torch_batch = ev, trn
        out = {}
        np.random.seed(c["rng_seed"]); torch.manual_seed(c["rng_seed"])
        for call, (seed, n) in enumerate(c["adds"]):
            pool.add_samples(cases.dyn_samples(seed, n))
            rec["eval"], rec["train"] = [], []
            stats = tr.train_with_pool(pool)
            pre = "c%d_" % call
            out[pre + "eval"] = np.stack(rec["eval"])
            out[pre + "train"] = np.stack(rec["train"])
            out[pre + "train_step"] = np.array(stats["train_step"])
            out[pre + "min_loss"] = np.array(tr.min_loss, dtype=np.float64)
            out[pre + "elite_indices"] = np.asarray(m.elite_indices, dtype=np.int64)
            out[pre + "rank"] = np.asarray(m.rank, dtype=np.int64)
            out[pre + "lengths"] = np.array([tr.train_length, tr.eval_length])
            out[pre + "index_proj"] = proj(T(tr.ensemble_index.astype(np.float64)), 5)
            for k, (mu, sd) in tr.mean_std_dict.items():
                out[pre + "mean_" + k], out[pre + "std_" + k] = np.asarray(mu), np.asarray(sd)
            nt = named_trainables(m)
            out[pre + "param_proj"] = np.stack([proj(p, 2000 + i) for i, (_, p) in enumerate(nt)])
            # how decisive every early-stopping comparison was (a fixture with near-ties would make
            # the trajectory depend on rounding): min |l - min_loss| / |l| over all decisions
            ev_arr = out[pre + "eval"][:-1]
            best = ev_arr[0].copy(); margin = np.inf
            for row in ev_arr[1:]:
                margin = min(margin, np.min(np.abs(row - best) / np.abs(row)))
                best = np.minimum(best, row)
            out[pre + "decision_margin"] = np.array(margin)
            print(tag, "call", call, "steps", stats["train_step"], "evals", len(rec["eval"]),
                  "margin %.2e" % margin,
                  "elites", m.elite_indices)
        out["names"] = np.array([n for n, _ in named_trainables(m)])
        np.savez(os.path.join(HERE, "pe_trainloop_%s.npz" % tag), **out)


def _margin(no, env):
    """Distance of every next_obs row to the nearest done threshold (hopper)."""
    assert env == "hopper"
    return np.minimum(np.abs(no[:, 0] - 0.7), np.abs(np.abs(no[:, 1]) - 0.2)).min()


def gen_collect():
    """MBPOCollector.imagine / MOPOCollector.imagine (mbpo_collector.py:34-97, mopo_collector.py:29-68)
    into a SimplePool: pool contents row for row.  The policy is tests.golden.cases.TanhPolicy (the
    reference GaussianPolicy is not on the path under test); its draws and the model's draws are
    recorded in consumption order."""
    from mirl.agents.mbpo.mbpo_collector import MBPOCollector
    from mirl.agents.mbpo.mopo_collector import MOPOCollector
    from mirl.pools.simple_pool import SimplePool
    from mirl.utils.logger import logger
    if hasattr(logger, "set_log_level"):
        logger.set_log_level(logger.ERROR)
    for tag, c in cases.COLLECT.items():
        best = None
        for rng_seed in range(100, 160):
            m, _ = build_pe(c["cfg"], c["mseed"], c["env"])
            m.remember_loss(c["loss"])
            env = ref_shim.FakeEnv(c["env"])
            pool = SimplePool(env, max_size=c["pool_n"] + 10)
            pool.add_samples(cases.dyn_samples(c["pool_seed"], c["pool_n"], env=c["env"]))
            imag = SimplePool(env, max_size=10)
            pol = cases.TanhPolicy(c["mseed"] + 5)
            kw = dict(depth_schedule=c["depth"], number_sample=c["number_sample"], save_n=c["save_n"],
                      bathch_size=c["batch_size"])
            if tag == "mopo":
                col = MOPOCollector(m, pol, pool, imag, penalty_coef=c["penalty_coef"],
                                    penalty_type=c["penalty_type"], **kw)
            else:
                col = MBPOCollector(m, pol, pool, imag, **kw)
            noise = []
            import mirl.agents.mbpo.pe_model as pem
            real_randn_like = torch.randn_like

            def rec_randn_like(t, *a, **k):
                e = real_randn_like(t, *a, **k)
                noise.append(e.numpy().copy())
                return e
            np.random.seed(rng_seed); torch.manual_seed(rng_seed)
            torch.randn_like = rec_randn_like
            try:
                col.imagine(0)
            finally:
                torch.randn_like = real_randn_like
            data = imag.get_data()
            mg = _margin(data["next_observations"], c["env"])
            if best is None or mg > best[0]:
                best = (mg, rng_seed, data, noise, pol.drawn, imag.max_size, imag.get_size())
            if mg > 1e-4:
                break
        mg, rng_seed, data, noise, pdraws, max_size, size = best
        print(tag, "seed", rng_seed, "rows", size, "margin %.2e" % mg, "steps", len(noise),
              "done ratio %.3f" % data["terminals"].mean())
        out = {"rng_seed": np.array(rng_seed), "margin": np.array(mg), "size": np.array([size, max_size]),
               "n_calls": np.array(len(noise)),
               "model_noise": np.concatenate(noise, 0), "policy_noise": np.concatenate(pdraws, 0),
               "call_rows": np.array([len(x) for x in noise])}
        for k, v in data.items():
            out["data_" + k] = v
        np.savez(os.path.join(HERE, "collector_%s.npz" % tag), **out)


def gen_extrapool():
    """ExtraFieldPool (extra_field_pool.py:19-71, flag_pool.py, utils/mean_std.py:12-33): running
    fp64 mean/std over three fills, the 'deltas' extra field, the compute_delta flag protocol of
    train_with_pool (pe_model_trainer.py:140-144), get_data order after a wrap, random batches."""
    from mirl.pools.extra_field_pool import ExtraFieldPool
    env = ref_shim.FakeEnv()
    pool = ExtraFieldPool(env, max_size=1000, compute_mean_std=True)
    pool.add_extra_fields({"deltas": {"shape": (cases.OBS,), "type": np.float32}})
    out = {}
    np.random.seed(93)
    for i, n in enumerate([400, 350, 500]):          # the third fill wraps
        pool.add_samples(cases.dyn_samples(9300 + i, n))
        tmp = pool.get_unprocessed_data("compute_delta", ["observations", "next_observations"])
        deltas = tmp["next_observations"] - tmp["observations"]
        out["unprocessed_%d" % i] = np.array(len(deltas))
        pool.update_single_extra_field("deltas", deltas)
        pool.update_process_flag("compute_delta", len(deltas))
        ms = pool.get_mean_std()
        for k, (mu, sd) in ms.items():
            out["mean_%d_%s" % (i, k)], out["std_%d_%s" % (i, k)] = np.asarray(mu), np.asarray(sd)
        b = pool.random_batch(32)
        for k, v in b.items():
            out["batch_%d_%s" % (i, k)] = v
        out["size_%d" % i] = np.array([pool.get_size(), pool._stop])
    for k, v in pool.get_data().items():
        out["data_" + k] = proj(T(v), 9)
    np.savez(os.path.join(HERE, "extra_pool.npz"), **out)


def gen_critic_update():
    """Critic half of SACAgent.train_from_torch_batch (sac_agent.py:160-166): DDPGAgent.compute_qf_loss
    (ddpg_agent.py:138-149) called on the reference class, torch.optim.Adam(qf.parameters(), lr=3e-4)
    (ddpg_agent.py:82-86) and ptu.soft_update_from_to(qf, qf_target, 5e-3) -- ten REDQ-shaped updates
    (10 critics, 2x256 relu, batch 256) plus one with clip_error."""
    import types
    import mirl.torch_modules.utils as ptu
    from mirl.agents.ddpg_agent import DDPGAgent
    from mirl.values.ensemble_value import EnsembleQValue
    env = ref_shim.FakeEnv()
    c = cases.REDQ
    sizes = [23] + c["hidden"] + [c["out"]]

    def build(seed):
        q = EnsembleQValue(env, ensemble_size=c["E"], hidden_layers=list(c["hidden"]), activation=c["act"])
        W, b = cases.mlp_params(seed, c["E"], sizes)
        q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W, b).items()})
        return q
    for tag, clip in (("plain", None), ("clip", 0.5)):
        qf, qt = build(71), build(171)
        opt = torch.optim.Adam(qf.parameters(), lr=3e-4)
        stub = types.SimpleNamespace(qf=qf, clip_error=clip, _log_q_info=lambda *a: {})
        out = {"loss": []}
        steps = 10 if clip is None else 3
        for s_ in range(steps):
            x = cases.critic_inputs(720 + s_, 256)
            q_target = T(cases.critic_q_target(x))
            loss, _ = DDPGAgent.compute_qf_loss(stub, T(x["obs"]), T(x["act"]), q_target)
            opt.zero_grad()
            loss.backward()
            opt.step()
            ptu.soft_update_from_to(qf, qt, 5e-3)
            out["loss"].append(float(loss))
            if s_ + 1 in (1, steps):
                names = [n for n, _ in qf.named_parameters()]
                out["param_proj_%d" % (s_ + 1)] = np.stack([proj(p, 5000 + i) for i, p in enumerate(qf.parameters())])
                out["target_proj_%d" % (s_ + 1)] = np.stack([proj(p, 6000 + i) for i, p in enumerate(qt.parameters())])
                out["m_proj_%d" % (s_ + 1)] = np.stack([proj(opt.state[p]["exp_avg"], 7000 + i)
                                                         for i, p in enumerate(qf.parameters())])
        out["loss"] = np.array(out["loss"], dtype=np.float64)
        out["names"] = np.array(names)
        np.savez(os.path.join(HERE, "critic_update_%s.npz" % tag), **out)
        print(tag, out["loss"])


def gen_tqc_update():
    """Critic half of TQCAgent's update (sac_agent.py:160-166 with tqc_agent.py:59-66): TQCAgent.compute_qf_loss called
    on the reference classes (TQCQValue 5 x [512,512,512] relu x 25 quantiles, configs/tqc.json), torch.optim.Adam
    (lr 3e-4) and ptu.soft_update_from_to(qf, qf_target, 5e-3): five updates on batch 256 against 115 target atoms."""
    import types
    import mirl.torch_modules.utils as ptu
    from mirl.agents.tqc_agent import TQCAgent
    from mirl.values.distributional_value import TQCQValue
    env = ref_shim.FakeEnv()
    c = cases.TQC
    sizes = [23] + c["hidden"] + [c["out"]]

    def build(seed):
        q = TQCQValue(env, ensemble_size=c["E"], number_head=c["out"], hidden_layers=list(c["hidden"]),
                      activation=c["act"])
        W, b = cases.mlp_params(seed, c["E"], sizes)
        q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W, b).items()})
        return q
    qf, qt = build(73), build(173)
    opt = torch.optim.Adam(qf.parameters(), lr=3e-4)
    stub = types.SimpleNamespace(qf=qf, _log_q_info=lambda *a: {})
    out = {"loss": []}
    steps = 5
    for s_ in range(steps):
        x = cases.critic_inputs(740 + s_, 256)
        atoms_target = T(cases.tqc_atoms_target(x, 760 + s_))
        loss, _ = TQCAgent.compute_qf_loss(stub, T(x["obs"]), T(x["act"]), atoms_target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ptu.soft_update_from_to(qf, qt, 5e-3)
        out["loss"].append(float(loss))
        if s_ + 1 in (1, steps):
            names = [n for n, _ in qf.named_parameters()]
            out["param_proj_%d" % (s_ + 1)] = np.stack([proj(p, 5000 + i) for i, p in enumerate(qf.parameters())])
            out["target_proj_%d" % (s_ + 1)] = np.stack([proj(p, 6000 + i) for i, p in enumerate(qt.parameters())])
            out["m_proj_%d" % (s_ + 1)] = np.stack([proj(opt.state[p]["exp_avg"], 7000 + i)
                                                     for i, p in enumerate(qf.parameters())])
    out["loss"] = np.array(out["loss"], dtype=np.float64)
    out["names"] = np.array(names)
    np.savez(os.path.join(HERE, "tqc_update.npz"), **out)
    print("tqc_update", out["loss"])


def gen_actor_update():
    """Actor half of SACAgent.train_from_torch_batch (sac_agent.py:170-195) on the reference classes: GaussianPolicy
    (17 -> 256 -> 256 -> 12 relu, tanh-squashed) and EnsembleQValue (REDQ: 10 critics, 2x256 relu; value = mean over
    all members, configs/redq.json:96-99): policy.action(reparameterize=True, return_log_prob=True),
    SACAgent.compute_alpha_loss + Adam(log_alpha), SACAgent.compute_policy_loss + Adam(policy), lr 3e-4,
    target_entropy -1 (redq.json:95); four updates on batch 256.  The second fixture uses a 2-member critic with
    mode 'min' (SAC's default current_v_pi_kwargs)."""
    import types
    from mirl.agents.sac_agent import SACAgent
    from mirl.policies.gaussian_policy import GaussianPolicy
    from mirl.values.ensemble_value import EnsembleQValue
    env = ref_shim.FakeEnv()
    for tag, E, v_pi in (("redq", 10, {"sample_number": None, "mode": "mean"}), ("sacmin", 2, {"sample_number": None, "mode": "min"})):
        pol = GaussianPolicy(env, hidden_layers=[256, 256], activation="relu")
        W, b = cases.mlp_params(95, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
        pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
        qf = EnsembleQValue(env, ensemble_size=E, hidden_layers=[256, 256], activation="relu")
        cW, cb = cases.mlp_params(96, E, [cases.OBS + cases.ACT, 256, 256, 1])
        qf.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(cW, cb).items()})
        log_alpha = torch.zeros(1, requires_grad=True)
        stub = types.SimpleNamespace(qf=qf, log_alpha=log_alpha, use_automatic_entropy_tuning=True, target_entropy=-1.0,
                                     _log_policy_info=lambda *a: {})
        stub._get_alpha = lambda: log_alpha.exp().detach()
        popt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        aopt = torch.optim.Adam([log_alpha], lr=3e-4)
        out = {k: [] for k in ("policy_loss", "q_pi", "entropy", "alpha_value", "alpha_loss", "log_alpha")}
        steps = 4
        for s_ in range(steps):
            o, _ = cases.rollout_inputs(970 + s_, 256)
            torch.manual_seed(980 + s_)
            new_action, info = pol.action(T(o), reparameterize=True, return_log_prob=True)
            alpha_loss, ainfo = SACAgent.compute_alpha_loss(stub, info)
            aopt.zero_grad()
            alpha_loss.backward()
            aopt.step()
            policy_loss, _ = SACAgent.compute_policy_loss(stub, T(o), new_action, info, v_pi)
            popt.zero_grad()
            policy_loss.backward()
            popt.step()
            with torch.no_grad():
                qv, _ = qf.value(T(o), new_action, **v_pi)
            out["policy_loss"].append(float(policy_loss)); out["q_pi"].append(float(qv.mean()))
            out["entropy"].append(float(-info["log_prob"].mean()))
            out["alpha_value"].append(ainfo["alpha/value"]); out["alpha_loss"].append(ainfo["alpha/loss"])
            out["log_alpha"].append(float(log_alpha))
            if s_ + 1 in (1, steps):
                out["param_proj_%d" % (s_ + 1)] = np.stack([proj(p, 8000 + i) for i, p in enumerate(pol.parameters())])
                out["m_proj_%d" % (s_ + 1)] = np.stack([proj(popt.state[p]["exp_avg"], 8100 + i)
                                                         for i, p in enumerate(pol.parameters())])
        out = {k: (np.array(v, dtype=np.float64) if isinstance(v, list) else v) for k, v in out.items()}
        out["names"] = np.array([n for n, _ in pol.named_parameters()])
        np.savez(os.path.join(HERE, "actor_update_%s.npz" % tag), **out)
        print(tag, out["policy_loss"], out["log_alpha"])


def gen_sac_update():
    """The UNMODIFIED reference agents end to end.
    redq: SACAgent.train_from_torch_batch (sac_agent.py:152-204) with the settings of configs/redq.json (10 critics
    2x256 relu, q_target = min over a batch-wise random pair, policy through the mean of all ten,
    policy_update_freq 20, target_entropy -1, lr 3e-4, tau 5e-3), 24 updates on batch 256.
    tqc: TQCAgent (tqc_agent.py:16-66) with configs/tqc.json (5 critics 3x512 relu x 25 quantiles, 8 % of the pooled
    target atoms dropped, quantile-Huber critic loss, policy through the mean of all atoms, every update), 6 updates.
    Recorded: critic loss of every update, the actor / alpha statistics of the updates that have them, parameter
    checksums of qf, qf_target, policy and log_alpha at the end.  RNG: np.random.seed(31) (member pairs),
    torch.manual_seed(32) (one [256, 6] standard-normal draw per policy.action call, in call order)."""
    import contextlib
    import tempfile
    from mirl.agents.sac_agent import SACAgent
    from mirl.agents.tqc_agent import TQCAgent
    from mirl.policies.gaussian_policy import GaussianPolicy
    from mirl.utils.logger import logger
    from mirl.values.distributional_value import TQCQValue
    from mirl.values.ensemble_value import EnsembleQValue
    with contextlib.redirect_stdout(sys.stderr):
        logger.set_snapshot_dir(tempfile.mkdtemp())
    env = ref_shim.FakeEnv()
    for tag in ("redq", "tqc"):
        c = cases.REDQ if tag == "redq" else cases.TQC
        pol = GaussianPolicy(env, hidden_layers=[256, 256], activation="relu")
        W, b = cases.mlp_params(95, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
        pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
        sizes = [cases.OBS + cases.ACT] + c["hidden"] + [c["out"]]

        def build(seed):
            if tag == "redq":
                q = EnsembleQValue(env, ensemble_size=c["E"], hidden_layers=list(c["hidden"]), activation=c["act"])
            else:
                q = TQCQValue(env, ensemble_size=c["E"], number_head=c["out"], hidden_layers=list(c["hidden"]),
                              activation=c["act"])
            cW, cb = cases.mlp_params(seed, c["E"], sizes)
            q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(cW, cb).items()})
            return q
        qf, qt = build(71), build(171)                # DDPGAgent.__init__ copies qf into qf_target (ddpg_agent.py:68)
        if tag == "redq":
            agent = SACAgent(env, pol, qf, qt, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                             use_automatic_entropy_tuning=True, alpha_if_not_automatic=0, policy_update_freq=20,
                             target_entropy=-1, current_v_pi_kwargs={"sample_number": None, "mode": "mean"},
                             next_v_pi_kwargs={"batchwise_sample": True}, tb_log_freq=-1)
            steps, freq = 24, 20
        else:
            agent = TQCAgent(env, pol, qf, qt, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                             use_automatic_entropy_tuning=True, alpha_if_not_automatic=0, tb_log_freq=-1)
            steps, freq = 6, 1
        np.random.seed(31)
        torch.manual_seed(32)
        out = {"qf_loss": [], "actor_steps": []}
        for k in ("policy/loss", "policy/q_pi", "policy/entropy", "alpha/value", "alpha/loss"):
            out[k.replace("/", "_")] = []
        for s_ in range(steps):
            x = cases.sac_batch(1200 + s_, 256)
            info = agent.train_from_torch_batch({k: T(v) for k, v in x.items()})
            out["qf_loss"].append(float(info["qf/loss"]))
            if s_ % freq == 0:
                out["actor_steps"].append(s_)
                for k in ("policy/loss", "policy/q_pi", "policy/entropy", "alpha/value", "alpha/loss"):
                    out[k.replace("/", "_")].append(float(info[k]))
        out = {k: np.array(v, dtype=np.float64) for k, v in out.items()}
        out["log_alpha"] = np.array(float(agent.log_alpha))
        out["qf_proj"] = np.stack([proj(p, 9000 + i) for i, p in enumerate(qf.parameters())])
        out["qt_proj"] = np.stack([proj(p, 9100 + i) for i, p in enumerate(qt.parameters())])
        out["pol_proj"] = np.stack([proj(p, 9200 + i) for i, p in enumerate(pol.parameters())])
        np.savez(os.path.join(HERE, "sac_update_%s.npz" % tag), **out)
        print("sac_update", tag, out["qf_loss"][:3], out["qf_loss"][-3:], out["policy_loss"], out["log_alpha"])


def gen_policy():
    """GaussianPolicy.action (gaussian_policy.py:35-47 -> gaussian.py:55-97,199-204), 17 -> 256 -> 256 -> 12 relu,
    tanh-squashed: reparameterised and plain samples with log-prob / mean / std, and the deterministic action.
    The fixture also records the standard-normal draw the reference consumed (replayed: torch.randn(B, A))."""
    from mirl.policies.gaussian_policy import GaussianPolicy
    env = ref_shim.FakeEnv()
    pol = GaussianPolicy(env, hidden_layers=[256, 256], activation="relu")
    W, b = cases.mlp_params(91, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
    B = 300
    o, _ = cases.rollout_inputs(92, B)
    out = {}
    with torch.no_grad():
        for tag, rep in (("rsample", True), ("sample", False)):
            torch.manual_seed(93)
            a, info = pol.action(T(o), return_log_prob=True, reparameterize=rep, return_mean_std=True)
            torch.manual_seed(93)
            eps = torch.randn(B, cases.ACT)
            assert torch.allclose(torch.tanh(info["mean"] + info["std"] * eps), a, atol=1e-6), tag
            out["action_" + tag], out["log_prob_" + tag] = a.numpy(), info["log_prob"].numpy()
            out["mean"], out["std"], out["eps"] = info["mean"].numpy(), info["std"].numpy(), eps.numpy()
        with pol.deterministic_(True):
            a, _ = pol.action(T(o))
        out["action_deterministic"] = a.numpy()
    out["keys"] = np.array(sorted(pol.state_dict().keys()))
    np.savez(os.path.join(HERE, "policy.npz"), **out)


def main():
    ref_shim.install()
    torch.set_num_threads(1)        # bitwise-stable reductions for the fixtures
    only = os.environ.get("GOLDEN_ONLY")
    if only:                        # e.g. GOLDEN_ONLY=pool regenerates one fixture
        globals()["gen_" + only]()
    else:
        gen_forward(); gen_step(); gen_dist(); gen_loss(); gen_train(); gen_critics(); gen_keys(); gen_pool()
        gen_trainloop(); gen_collect(); gen_extrapool(); gen_critic_update(); gen_tqc_update(); gen_actor_update(); gen_sac_update(); gen_policy()
    for f in sorted(os.listdir(HERE)):
        if f.endswith((".npz", ".json")):
            print("%8d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))


if __name__ == "__main__":
    main()
