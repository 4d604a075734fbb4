"""Round-2 GPU parity tests (through the C ABI): the persistent fused training kernel, the training loop
against the reference's ``train_with_pool`` fixture, the collectors against the reference collectors'
pool contents, the device ``ExtraFieldPool``, tcgen05 at every ensemble size, and an oracle comparison
at 1e5 rows."""
import numpy as np
import pytest
import torch

from oracle import pe_oracle as po
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, done_guard_mismatches, make_pe_model
from tests.util import RTOL, T, assert_close, golden, norm_rel, proj, to_torch

pytestmark = pytest.mark.gpu
LOSS_VEC7 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4]


def _batch(seed, E, B, device="cuda"):
    return {k: T(v, device) for k, v in cases.train_batch(seed, E, B).items()}


def _trainer(m, **kw):
    import mirl_b200
    kw.setdefault("lr", 1e-3)
    return mirl_b200.B200PEModelTrainer(FakeEnv(), m, tb_log_freq=-1, silent=True, **kw)


# ---- fused persistent training kernel ------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 5, 16, 100, 256, 257, 600])
def test_fused_train_step_matches_oracle_ragged(B):
    """One step of the persistent kernel (train_fused.cu) against the CPU oracle: loss, per-member
    loss, parameters and both Adam moments -- batches that are not multiples of the 16-row block or
    of the 64-row chunk, and more row blocks than CTAs per member (B = 600)."""
    m, p = make_pe_model(cases.MBPO, 11)
    tr = _trainer(m, ignore_terminal_state=(B % 2 == 1))
    bt = _batch(900 + B, 7, B)
    diag = tr.train_from_torch_batch(bt)
    # (to_torch shares memory with the numpy parameters and the oracle's step updates in place: take the
    # analytic gradients of the ORIGINAL parameters first)
    g = po.model_loss_grads(to_torch(p), *[bt[k].cpu() for k in ("observations", "actions", "deltas", "rewards",
                                                                 "terminals")],
                            ignore_terminal_state=(B % 2 == 1))
    pt = to_torch(p)
    opt = po.new_adam_state(pt)
    loss, el = po.train_step(pt, opt, {k: v.cpu() for k, v in bt.items()},
                             ignore_terminal_state=(B % 2 == 1))
    assert_close(np.array(diag["train_loss"]), loss.numpy(), what="loss")
    assert_close(np.array([diag["mt/net_%d/train_loss" % i] for i in range(7)]), el.numpy(), what="ensemble_loss")
    sd = m.module.reference_state_dict("module.")
    msd = m.module.reference_state_dict("module.", flat=tr.exp_avg)
    pn = {}
    for k, v in pt.items():                                   # oracle parameters after its own step
        pn[k] = [t.numpy() for t in v] if isinstance(v, list) else (v.numpy() if torch.is_tensor(v) else v)
    ref = cases.pe_state_dict(pn)
    worst = 0.0
    for k, v in sd.items():
        # first Adam step is lr * g/(|g| + eps): sign-like, so tiny gradients may flip -- compare the
        # update through the first moment (the gradient itself) instead of element-wise parameters
        worst = max(worst, float(np.abs(v.cpu().numpy() - ref[k]).max()))
    assert worst <= 2.0e-3 + 1e-9, worst                      # |update| <= lr per element, both sides
    # exp_avg after one step = 0.1 * grad: norm-wise 1e-3 against the oracle's analytic gradients
    for l in range(5):
        for e in range(7):
            got = msd["module.net.%d.weight_net%d" % (2 * l, e)].cpu().numpy() / 0.1
            assert norm_rel(got, g["W"][l][e].numpy()) < 1e-3, ("dW", l, e)
            gb = msd["module.net.%d.bias_net%d" % (2 * l, e)].cpu().numpy() / 0.1
            assert norm_rel(gb, g["b"][l][e].numpy()) < 1e-3, ("db", l, e)
    for e in range(7):
        assert norm_rel(msd["module.maxlogstd_net%d" % e].cpu().numpy() / 0.1, g["max_logstd"][e].numpy()) < 1e-3
        assert norm_rel(msd["module.minlogstd_net%d" % e].cpu().numpy() / 0.1, g["min_logstd"][e].numpy()) < 1e-3


@pytest.mark.parametrize("E,hidden,B", [(2, [200, 200, 200, 200], 256), (10, [200, 200, 200, 200], 256),
                                        (20, [64, 48], 300), (3, [256, 256], 100)])
def test_fused_train_step_ensemble_sizes_and_widths(E, hidden, B):
    """The persistent kernel at other ensemble sizes and layer widths: 2 members (21 CTAs each, 5 of them without a
    row block), 10 members (MOPO: 14 CTAs per member, two row blocks on some), 20 members (7 CTAs per member, three
    row blocks each, narrow layers with partial tiles) and 256-wide layers (the 16 k-step instantiation): loss,
    per-member loss and the first Adam moments against the oracle's analytic gradients."""
    cfg = dict(E=E, elites=max(1, E - 1), hidden=hidden, act="swish")
    m, p = make_pe_model(cfg, 23)
    tr = _trainer(m)
    assert tr.fused_steps_supported()
    bt = _batch(950 + E, E, B)
    diag = tr.train_from_torch_batch(bt)
    g = po.model_loss_grads(to_torch(p), *[bt[k].cpu() for k in ("observations", "actions", "deltas", "rewards",
                                                                 "terminals")],
                            weight_decay=cases.WEIGHT_DECAY[:len(hidden) + 1])
    assert_close(np.array(diag["train_loss"]), g["loss"].numpy(), what="loss")
    assert_close(np.array([diag["mt/net_%d/train_loss" % i] for i in range(E)]), g["ensemble_loss"].numpy(),
                 what="ensemble_loss")
    msd = m.module.reference_state_dict("module.", flat=tr.exp_avg)
    for l in range(len(hidden) + 1):
        for e in range(E):
            got = msd["module.net.%d.weight_net%d" % (2 * l, e)].cpu().numpy() / 0.1
            assert norm_rel(got, g["W"][l][e].numpy()) < 1e-3, ("dW", l, e)
            gb = msd["module.net.%d.bias_net%d" % (2 * l, e)].cpu().numpy() / 0.1
            assert norm_rel(gb, g["b"][l][e].numpy()) < 1e-3, ("db", l, e)
    for e in range(E):
        assert norm_rel(msd["module.maxlogstd_net%d" % e].cpu().numpy() / 0.1, g["max_logstd"][e].numpy()) < 1e-3
        assert norm_rel(msd["module.minlogstd_net%d" % e].cpu().numpy() / 0.1, g["min_logstd"][e].numpy()) < 1e-3


def test_fused_train_100_steps_match_reference_adam():
    """A hundred fused steps against the reference's PEModelTrainer run (loss curve, parameter and
    moment checksums at step 100)."""
    g = golden("pe_train_mbpo100.npz")
    m, _ = make_pe_model(cases.MBPO, 11)
    tr = _trainer(m, ignore_terminal_state=False)
    for s in range(100):
        el, terms = tr.train_step_async(_batch(600 + s, 7, 256))
        if s % 10 == 9 or s < 3:
            assert_close(terms.sum().cpu().numpy(), g["loss"][s], what="loss step %d" % s)
            assert_close(el.cpu().numpy(), g["ensemble_loss"][s], what="ensemble_loss step %d" % s)
    sd = m.state_dict()
    msd = m.module.reference_state_dict("module.", flat=tr.exp_avg)
    vsd = m.module.reference_state_dict("module.", flat=tr.exp_avg_sq)
    for i, name in enumerate(g["names"]):
        name = str(name)
        got, ref = proj(sd[name], 2000 + i), g["param_proj_100"][i]
        tol = 1e-3 * 100 * 0.02 * np.sqrt(sd[name].numel()) + 1e-4 * ref[0]
        assert abs(got[0] - ref[0]) <= tol, (name, got, ref)
        gm, rm = proj(msd[name], 3000 + i), g["m_proj_100"][i]
        assert abs(gm[0] - rm[0]) <= 5e-3 * rm[0] + 1e-9, ("exp_avg", name, gm, rm)
        gv, rv = proj(vsd[name], 4000 + i), g["v_proj_100"][i]
        assert abs(gv[0] - rv[0]) <= 5e-3 * rv[0] + 1e-12, ("exp_avg_sq", name, gv, rv)


def test_multi_step_launch_equals_single_step_launches():
    """mb_pe_train_steps (bootstrap batches gathered on the device from the index table, ragged last
    batch, epoch wrap-around) is bit-identical to one launch per step on host-gathered batches."""
    cfg = cases.MBPO
    N, bs, nsteps = 700, 256, 7                     # 3 steps per epoch: 256, 256, 188
    data = {k: T(v, "cuda") for k, v in cases.train_batch(77, 1, N, shared=True).items()}
    rs = np.random.RandomState(5)
    index = torch.from_numpy(rs.randint(N, size=(cfg["E"], N))).cuda()
    ma, _ = make_pe_model(cfg, 11)
    mb_, _ = make_pe_model(cfg, 11)
    ta, tb = _trainer(ma), _trainer(mb_)
    el, terms = ta.train_steps_async(data, index, N, bs, 0, nsteps)
    spe = (N + bs - 1) // bs
    for t in range(nsteps):
        j = (t % spe) * bs
        ind = index[:, j:j + bs]
        e1, t1 = tb.train_step_async({k: v[ind] for k, v in data.items()})
        assert torch.equal(e1, el[t]), ("ensemble_loss", t)
        assert_close(t1, terms[t], 1e-5, "loss terms %d" % t)
    torch.cuda.synchronize()
    assert ta.num_train_steps == tb.num_train_steps == nsteps
    assert torch.equal(ma.module.flat.data, mb_.module.flat.data)
    assert torch.equal(ta.exp_avg, tb.exp_avg) and torch.equal(ta.exp_avg_sq, tb.exp_avg_sq)
    # continuing from step 7 in a second launch == one launch of 10
    mc, _ = make_pe_model(cfg, 11)
    tc = _trainer(mc)
    tc.train_steps_async(data, index, N, bs, 0, 10)
    ta.train_steps_async(data, index, N, bs, 7, 3)
    assert torch.equal(ma.module.flat.data, mc.module.flat.data)


def test_member_sharded_step_equals_unsharded_bitwise():
    """Member-sharded training (last matrix row): two slices trained by two trainers over the
    same blob equal the unsharded step bit for bit (members are independent, pe_model.py:95-96)."""
    m1, _ = make_pe_model(cases.MBPO, 11)
    m2, _ = make_pe_model(cases.MBPO, 11)
    full = _trainer(m1)
    lo, hi = _trainer(m2, members=(0, 4)), _trainer(m2, members=(4, 7))
    hi.exp_avg, hi.exp_avg_sq = lo.exp_avg, lo.exp_avg_sq           # one optimiser state, two slices
    for s in range(3):
        bt = _batch(600 + s, 7, 256)
        e_full, _ = full.train_step_async(bt)
        e_full = e_full.clone()
        e_lo, _ = lo.train_step_async(bt)
        e_lo = e_lo.clone()
        e_hi, _ = hi.train_step_async(bt)
        assert torch.equal(e_full[:4], e_lo[:4]) and torch.equal(e_full[4:], e_hi[4:])
    assert torch.equal(m1.module.flat.data, m2.module.flat.data)
    assert torch.equal(full.exp_avg, lo.exp_avg)
    # an empty slice (E=7 over 8 ranks) is a no-op
    none = _trainer(m2, members=(7, 7))
    before = m2.module.flat.data.clone()
    none.train_steps_async({k: v[0] for k, v in bt.items()}, torch.zeros(7, 256, dtype=torch.int64, device="cuda"),
                           256, 256, 0, 2)
    assert torch.equal(before, m2.module.flat.data)


def test_train_b1_ant_sized_model_chain_path(monkeypatch):
    """B = 1 with a wide D (AntTruncatedObs: O=27, A=8) -- the chain path's prep kernel must
    zero all 2*E*D bound gradients even when its grid has fewer threads than that."""
    import ctypes
    from mirl_b200 import _lib
    m, p = make_pe_model(cases.MBPO, 11, env="mb_ant", obs=27, act=8)
    L = _lib.lib()
    bt = {k: T(v, "cuda") for k, v in cases.train_batch(3, 7, 1, obs=27, act=8).items()}
    cm = m._c_model(None)
    ws = torch.empty(L.mb_pe_train_workspace_bytes(ctypes.byref(cm), 1), dtype=torch.uint8, device="cuda")
    cfg = m.train_cfg(False)
    pt = to_torch(p)
    want = po.model_loss_grads(pt, *[bt[k].cpu() for k in ("observations", "actions", "deltas", "rewards", "terminals")])
    for rep in range(2):                         # the second call would see the first call's leftovers
        grads = torch.full_like(m.module.flat.data, 7.0)
        el, terms = torch.zeros(7, device="cuda"), torch.zeros(3, device="cuda")
        _lib.check(L.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(cfg), _lib.ptr(bt["observations"]),
                                      _lib.ptr(bt["actions"]), _lib.ptr(bt["deltas"]), _lib.ptr(bt["rewards"]),
                                      _lib.ptr(bt["terminals"]), 1, _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms),
                                      _lib.ptr(ws), ws.numel(), _lib.stream()))
        gsd = m.module.reference_state_dict("module.", flat=grads)
        for e in range(7):
            assert norm_rel(gsd["module.maxlogstd_net%d" % e].cpu().numpy(), want["max_logstd"][e].numpy()) < 1e-3
            assert norm_rel(gsd["module.minlogstd_net%d" % e].cpu().numpy(), want["min_logstd"][e].numpy()) < 1e-3


# ---- the training loop against the reference's train_with_pool -----------------------------------
@pytest.mark.parametrize("tag", ["small", "mbpo"])
def test_train_with_pool_matches_reference_fixture(tag):
    """B200PEModelTrainer.train_with_pool over a B200ExtraFieldPool against the UNMODIFIED reference
    (tests/golden/gen_golden.py::gen_trainloop): normaliser statistics, split lengths, bootstrap table,
    hold-out loss at every report point, per-step training loss, early-stopping step count, restored
    best snapshots (parameter checksums) and the final elite ranking -- two consecutive calls."""
    import mirl_b200
    from mirl_b200.pools import B200ExtraFieldPool
    c = cases.TRAINLOOP[tag]
    cfg = c["cfg"]
    g = golden("pe_trainloop_%s.npz" % tag)
    m, _ = make_pe_model(cfg, c["seed"])
    env = FakeEnv()
    pool = B200ExtraFieldPool(env, max_size=5000, compute_mean_std=True)
    pool.add_extra_fields({"deltas": {"shape": (cases.OBS,), "type": np.float32}})
    tr = mirl_b200.B200PEModelTrainer(env, m, lr=c["lr"], tb_log_freq=-1, silent=True, ignore_terminal_state=False,
                                      batch_size=c["batch_size"], report_freq=c["report_freq"],
                                      max_not_improve=c["max_not_improve"], max_valid=c["max_valid"],
                                      init_model_train_step=c["init_model_train_step"],
                                      max_step_per_train=c["max_step_per_train"])
    rec = {"eval": [], "train": []}
    ev0, st0 = tr.eval_with_dataset, tr.train_steps_async

    def ev(ds):
        s, el = ev0(ds)
        rec["eval"].append(np.array(el, dtype=np.float64))
        return s, el

    def st(*a, **k):
        el, terms = st0(*a, **k)
        rec["train"].append(torch.cat([el, terms.sum(1, keepdim=True)], 1).cpu().numpy().astype(np.float64))
        return el, terms
    tr.eval_with_dataset, tr.train_steps_async = ev, st
    # lr = 0.1 on the small config is a deliberately rough ride (early stopping must trigger): rounding
    # differences are amplified step by step, the DECISIONS (margin >= 2e-2 in the fixture) must not move
    tol = 5e-2 if tag == "small" else 2e-3
    np.random.seed(c["rng_seed"])
    torch.manual_seed(c["rng_seed"])
    for call, (seed, n) in enumerate(c["adds"]):
        pool.add_samples(cases.dyn_samples(seed, n))
        rec["eval"], rec["train"] = [], []
        stats = tr.train_with_pool(pool)
        pre = "c%d_" % call
        assert float(g[pre + "decision_margin"]) > 1e-3          # no early-stopping decision is a near-tie
        assert [tr.train_length, tr.eval_length] == list(g[pre + "lengths"])
        assert np.allclose(proj(torch.as_tensor(tr.ensemble_index.astype(np.float64)), 5), g[pre + "index_proj"])
        for k, (mu, sd) in tr.mean_std_dict.items():
            assert np.allclose(mu, g[pre + "mean_" + k], rtol=1e-5, atol=1e-7), k
            assert np.allclose(sd, g[pre + "std_" + k], rtol=1e-5, atol=1e-7), k
        assert stats["train_step"] == int(g[pre + "train_step"])
        got_eval = np.stack(rec["eval"])
        assert got_eval.shape == g[pre + "eval"].shape
        assert np.abs(got_eval - g[pre + "eval"]).max() <= tol * np.abs(g[pre + "eval"]).max(), \
            np.abs(got_eval / g[pre + "eval"] - 1).max()
        assert np.allclose(got_eval, g[pre + "eval"], rtol=tol * 4, atol=0)
        got_train = np.concatenate(rec["train"])
        assert got_train.shape == g[pre + "train"].shape
        assert np.allclose(got_train, g[pre + "train"], rtol=tol * 4, atol=0)
        assert list(m.elite_indices) == list(g[pre + "elite_indices"])
        assert list(m.rank) == list(g[pre + "rank"])
        assert np.allclose(np.array(tr.min_loss, dtype=np.float64), g[pre + "min_loss"], rtol=tol * 4)
        sd = m.state_dict()
        for i, name in enumerate(g["names"]):
            got, ref = proj(sd[str(name)], 2000 + i), g[pre + "param_proj"][i]
            assert abs(got[0] - ref[0]) <= (tol * 2) * ref[0] + 2e-3 * np.sqrt(sd[str(name)].numel()) * c["lr"] * 50, \
                (str(name), got, ref)


def test_extra_field_pool_matches_reference():
    """B200ExtraFieldPool against the reference ExtraFieldPool: float64 running mean/std over three fills
    (the third wraps the ring), the deltas field, the compute_delta flag protocol, sampled batches."""
    from mirl_b200.pools import B200ExtraFieldPool
    g = golden("extra_pool.npz")
    pool = B200ExtraFieldPool(FakeEnv(), max_size=1000, compute_mean_std=True)
    pool.add_extra_fields({"deltas": {"shape": (cases.OBS,), "type": np.float32}})
    np.random.seed(93)
    for i, n in enumerate([400, 350, 500]):
        pool.add_samples(cases.dyn_samples(9300 + i, n))
        tmp = pool.get_unprocessed_data_torch("compute_delta", ["observations", "next_observations"])
        deltas = tmp["next_observations"] - tmp["observations"]
        assert len(deltas) == int(g["unprocessed_%d" % i])
        with pytest.raises(RuntimeError):
            pool.random_batch(4)                               # deltas not aligned yet
        pool.update_single_extra_field("deltas", deltas)
        pool.update_process_flag("compute_delta", len(deltas))
        for k, (mu, sd) in pool.get_mean_std().items():
            assert np.allclose(mu, g["mean_%d_%s" % (i, k)], rtol=1e-5, atol=1e-6), (i, k)
            assert np.allclose(sd, g["std_%d_%s" % (i, k)], rtol=1e-5, atol=1e-6), (i, k)
        b = pool.random_batch(32)
        for k, v in b.items():
            assert np.array_equal(v, g["batch_%d_%s" % (i, k)]), (i, k)
        assert [pool.get_size(), pool._stop] == list(g["size_%d" % i])
    for k, v in pool.get_data().items():
        assert np.allclose(proj(torch.from_numpy(v), 9), g["data_" + k]), k


# ---- collectors against the reference collectors --------------------------------------------------
class _ReplayModel:
    """Feeds the recorded torch.randn draws of the reference run to the model (the CUDA generator is a
    different stream); member indices still come from the global NumPy generator, like the reference."""

    def __init__(self, model, noise):
        self._m, self._noise, self._pos = model, noise, 0

    def __getattr__(self, k):
        return getattr(self._m, k)

    def _take(self, n):
        out = T(self._noise[self._pos:self._pos + n], "cuda")
        self._pos += n
        return out

    def step(self, o, a, **kw):
        return self._m.step(o, a, noise=self._take(o.shape[0]), **kw)

    def step_rows(self, o, a, dest, **kw):
        return self._m.step_rows(o, a, dest, noise=self._take(o.shape[0]), **kw)


class _Replay:
    def __init__(self, arr):
        self.arr, self.pos = arr, 0

    def __call__(self, B, A):
        out = T(self.arr[self.pos:self.pos + B], "cuda")
        self.pos += B
        return out


@pytest.mark.parametrize("tag,imagined", [("mbpo", "device"), ("mbpo", "host"), ("mopo", "host")])
def test_collector_matches_reference_pool_contents(tag, imagined):
    """B200MBPOCollector / B200MOPOCollector.imagine against the pool the reference collector filled
    (hopper, two chunks, depth 5 / 3, rows terminate and are compacted away): same NumPy seed, recorded
    torch draws -> same start states, same member indices, and the pool contents match row for row."""
    import mirl_b200
    from mirl_b200.collectors import B200MBPOCollector, B200MOPOCollector
    c = cases.COLLECT[tag]
    g = golden("collector_%s.npz" % tag)
    env = FakeEnv(c["env"])
    m, _ = make_pe_model(c["cfg"], c["mseed"], c["env"])
    m.remember_loss(c["loss"])
    pool = mirl_b200.B200DevicePool(env, max_size=c["pool_n"] + 10)
    pool.add_samples(cases.dyn_samples(c["pool_seed"], c["pool_n"], env=c["env"]))
    imag = mirl_b200.B200DevicePool(env, max_size=10, row_stride=m.row_stride())
    pol = cases.TanhPolicy(c["mseed"] + 5, noise_source=_Replay(g["policy_noise"]))
    model = _ReplayModel(m, g["model_noise"])
    kw = dict(depth_schedule=c["depth"], number_sample=c["number_sample"], save_n=c["save_n"],
              bathch_size=c["batch_size"])
    if tag == "mopo":
        col = B200MOPOCollector(model, pol, pool, imag, penalty_coef=c["penalty_coef"],
                                penalty_type=c["penalty_type"], **kw)
    else:
        col = B200MBPOCollector(model, pol, pool, imag, **kw)
    if imagined == "host":
        col._fused_step = lambda o, a: None                  # reference-shaped path: step + add_samples
    np.random.seed(int(g["rng_seed"]))
    torch.manual_seed(int(g["rng_seed"]))
    col.imagine(0)
    assert [imag.get_size(), imag.max_size] == list(g["size"])
    assert model._pos == len(g["model_noise"]) and pol.noise_source.pos == len(g["policy_noise"])
    data = imag.get_data()
    assert np.array_equal(data["terminals"], g["data_terminals"])
    for k in ("observations", "actions", "next_observations", "rewards"):
        assert_close(data[k], g["data_" + k], what=k)
    assert 0.05 < data["terminals"].mean() < 0.95


# ---- tensor-core kernel at every ensemble size; oracle comparison at 1e5 rows ----------------------
@pytest.mark.parametrize("E,elites", [(2, 2), (5, 3), (10, 7), (20, 15)])
def test_tcgen05_ensemble_sizes_match_oracle(E, elites):
    cfg = dict(E=E, elites=elites, hidden=[200, 200, 200, 200], act="swish")
    m, p = make_pe_model(cfg, 60 + E, "walker2d", kernel="tcgen05")
    if not m._tc_supported():
        pytest.skip("tcgen05 kernel does not cover this shape")
    loss = list(np.random.RandomState(E).uniform(size=E))
    m.remember_loss(loss)
    B = 777
    o, a = cases.rollout_inputs(70 + E, B, env="walker2d")
    eps = np.random.RandomState(71).standard_normal((B, 18)).astype(np.float32)
    idx = np.random.RandomState(72).choice(m.elite_indices, B)
    no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), indices=idx, noise=T(eps, "cuda"))
    rno, rr, rd = po.rollout_step(to_torch(p), T(o), T(a), torch.from_numpy(idx), T(eps), "walker2d")
    assert_close(no, rno, what="next_obs")
    assert_close(r, rr, what="reward")
    out, _ = done_guard_mismatches("walker2d", rno.numpy(), rd.numpy(), d.cpu().numpy())
    assert out == 0
    mean, ls = m.predict_distribution(T(o[:300], "cuda"), T(a[:300], "cuda"))
    rmean, rls = po.ensemble_distribution(to_torch(p), T(o[:300]), T(a[:300]))
    assert_close(mean, rmean, what="mean")
    assert_close(ls, rls, what="logstd")


@pytest.mark.parametrize("kernel", ["tcgen05", "simt"])
def test_rollout_100k_rows_matches_oracle(kernel):
    """The bench's own shape (7 members / 5 elites, 4x200, obs 17 / act 6) at 1e5 rows against the CPU
    oracle, device-side member draw included."""
    B = 100_000
    m, p = make_pe_model(cases.MBPO, 11, "hopper", kernel=kernel)
    m.remember_loss(LOSS_VEC7)
    o, a = cases.rollout_inputs(5, B, env="hopper")
    eps = np.random.RandomState(6).standard_normal((B, 18)).astype(np.float32)
    np.random.seed(9)
    want_idx = np.random.choice(m.elite_indices, B)
    np.random.seed(9)
    no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), noise=T(eps, "cuda"))        # draws on the device
    assert m._drew_on_device
    torch.set_num_threads(max(1, torch.get_num_threads()))
    rno, rr, rd = po.rollout_step(to_torch(p), T(o), T(a), torch.from_numpy(want_idx), T(eps), "hopper")
    assert_close(no, rno, what="next_obs")
    assert_close(r, rr, what="reward")
    out, inside = done_guard_mismatches("hopper", rno.numpy(), rd.numpy(), d.cpu().numpy())
    assert out == 0 and inside <= 20


# ---- N3: fused critic update ------------------------------------------------------------------------
@pytest.mark.parametrize("tag,clip,steps", [("plain", None, 10), ("clip", 0.5, 3)])
def test_fused_critic_update_matches_reference(tag, clip, steps):
    """B200CriticTrainer against the reference's DDPGAgent.compute_qf_loss + torch.optim.Adam +
    soft_update_from_to run (REDQ shape: 10 critics, 2x256 relu, batch 256): loss curve, parameters, first
    moments and the soft-updated target network."""
    import mirl_b200
    from tests.helpers_gpu import make_critic
    g = golden("critic_update_%s.npz" % tag)
    qf, _ = make_critic("redq", 71)
    qt, _ = make_critic("redq", 171)
    tr = mirl_b200.B200CriticTrainer(qf, qt, qf_lr=3e-4, soft_target_tau=5e-3, clip_error=clip)
    assert tr.supported()
    for s_ in range(steps):
        x = cases.critic_inputs(720 + s_, 256)
        info = tr.update(T(x["obs"], "cuda"), T(x["act"], "cuda"), T(cases.critic_q_target(x), "cuda"))
        assert abs(info["qf/loss"] - g["loss"][s_]) <= 1e-3 * abs(g["loss"][s_]), (s_, info, g["loss"][s_])
        if s_ + 1 in (1, steps):
            msd = qf.module.reference_state_dict("module.", flat=tr.exp_avg)
            for i, (name, p, pt) in enumerate(zip(g["names"], qf.parameters(), qt.parameters())):
                name = str(name)
                n = p.numel()
                got, ref = proj(p, 5000 + i), g["param_proj_%d" % (s_ + 1)][i]
                assert abs(got[0] - ref[0]) <= 3e-4 * (s_ + 1) * 0.05 * np.sqrt(n) + 1e-4 * ref[0], (name, got, ref)
                got, ref = proj(pt, 6000 + i), g["target_proj_%d" % (s_ + 1)][i]
                assert abs(got[0] - ref[0]) <= 1e-4 * ref[0] + 1e-6, ("target", name, got, ref)
                assert abs(got[1] - ref[1]) <= 1e-3 * ref[0] + 1e-6, ("target", name, got, ref)
                gm, rm = proj(msd[name], 7000 + i), g["m_proj_%d" % (s_ + 1)][i]
                # (a [1,1] bias moment is one signed sum over the batch: cancellation, so an absolute floor)
                assert abs(gm[0] - rm[0]) <= 2e-3 * rm[0] + 2e-5, ("exp_avg", name, gm, rm)


def test_tqc_critic_update_matches_reference():
    """B200CriticTrainer on B200TQCQValue critics against the reference's TQCAgent.compute_qf_loss
    (quantile_huber_loss_f) + torch.optim.Adam + soft_update_from_to run (configs/tqc.json shape: 5 critics,
    3x512 relu, 25 quantiles, 115 target atoms, batch 256): loss curve, parameters, moments, target network."""
    import mirl_b200
    from tests.helpers_gpu import make_critic
    g = golden("tqc_update.npz")
    qf, _ = make_critic("tqc", 73)
    qt, _ = make_critic("tqc", 173)
    tr = mirl_b200.B200CriticTrainer(qf, qt, qf_lr=3e-4, soft_target_tau=5e-3)
    assert tr.supported()
    steps = len(g["loss"])
    for s_ in range(steps):
        x = cases.critic_inputs(740 + s_, 256)
        info = tr.update(T(x["obs"], "cuda"), T(x["act"], "cuda"), T(cases.tqc_atoms_target(x, 760 + s_), "cuda"))
        assert abs(info["qf/loss"] - g["loss"][s_]) <= 1e-3 * abs(g["loss"][s_]), (s_, info, g["loss"][s_])
        if s_ + 1 in (1, steps):
            msd = qf.module.reference_state_dict("module.", flat=tr.exp_avg)
            for i, (name, p, pt) in enumerate(zip(g["names"], qf.parameters(), qt.parameters())):
                name = str(name)
                n = p.numel()
                got, ref = proj(p, 5000 + i), g["param_proj_%d" % (s_ + 1)][i]
                assert abs(got[0] - ref[0]) <= 3e-4 * (s_ + 1) * 0.05 * np.sqrt(n) + 1e-4 * ref[0], (name, got, ref)
                got, ref = proj(pt, 6000 + i), g["target_proj_%d" % (s_ + 1)][i]
                assert abs(got[0] - ref[0]) <= 1e-4 * ref[0] + 1e-6, ("target", name, got, ref)
                assert abs(got[1] - ref[1]) <= 1e-3 * ref[0] + 1e-6, ("target", name, got, ref)
                gm, rm = proj(msd[name], 7000 + i), g["m_proj_%d" % (s_ + 1)][i]
                assert abs(gm[0] - rm[0]) <= 2e-3 * rm[0] + 2e-5, ("exp_avg", name, gm, rm)


def test_wide_critic_update_chain_matches_oracle():
    """mb_critic_train_step on critics wider than the persistent kernel covers (3 members, 2x384 relu,
    batch 200, clip_error): the launch-chain path against critic_oracle.critic_update (pinned to the reference by
    the critic_update fixtures), three steps with the soft target update."""
    import mirl_b200
    from oracle import critic_oracle as co
    E, sizes = 3, [23, 384, 384, 1]
    W, b = cases.mlp_params(75, E, sizes)
    tW, tb = cases.mlp_params(175, E, sizes)

    def build(W_, b_):
        q = mirl_b200.B200EnsembleQValue(FakeEnv(), ensemble_size=E, hidden_layers=sizes[1:-1], activation="relu")
        q.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(W_, b_).items()})
        return q.to("cuda")
    qf, qt = build(W, b), build(tW, tb)
    tr = mirl_b200.B200CriticTrainer(qf, qt, qf_lr=1e-3, soft_target_tau=0.01, clip_error=0.7)
    oW, ob = [T(a.copy()) for a in W], [T(a.copy()) for a in b]
    otW, otb = [T(a.copy()) for a in tW], [T(a.copy()) for a in tb]
    opt = {"step": 0, "m": [torch.zeros_like(t) for t in oW + ob], "v": [torch.zeros_like(t) for t in oW + ob]}
    rs = np.random.RandomState(5)
    for s_ in range(3):
        x = cases.critic_inputs(770 + s_, 200)
        tgt = rs.standard_normal((200, 1)).astype(np.float32)
        want = co.critic_update(oW, ob, opt, otW, otb, T(x["obs"]), T(x["act"]), T(tgt), lr=1e-3, tau=0.01,
                                clip_error=0.7)
        info = tr.update(T(x["obs"], "cuda"), T(x["act"], "cuda"), T(tgt, "cuda"))
        assert abs(info["qf/loss"] - float(want)) <= 1e-3 * abs(float(want)), (s_, info, float(want))
    sd, tsd = qf.module.reference_state_dict("module."), qt.module.reference_state_dict("module.")
    for l in range(len(oW)):
        for e in range(E):
            for key, o_, ot_ in (("weight", oW[l][e], otW[l][e]), ("bias", ob[l][e], otb[l][e])):
                name = "module.net.%d.%s_net%d" % (2 * l, key, e)
                # Adam's first steps move every weight by about lr whatever the gradient's size: compare to that scale
                assert (sd[name].cpu() - o_).abs().max() <= 0.15 * 3 * 1e-3, name
                assert_close(tsd[name].cpu(), ot_, what="target " + name)


@pytest.mark.parametrize("tag,E,mode", [("redq", 10, "mean"), ("sacmin", 2, "min")])
def test_actor_update_matches_reference(tag, E, mode):
    """B200ActorTrainer against the reference's actor half of SACAgent.train_from_torch_batch (policy.action with the
    recorded rsample draw, compute_alpha_loss + Adam(log_alpha), compute_policy_loss + Adam(policy); REDQ: mean over
    10 critics, SAC default: min over 2): policy loss, q_pi, entropy, alpha trajectory, policy parameters and first
    moments after 1 and 4 updates."""
    import mirl_b200
    g = golden("actor_update_%s.npz" % tag)
    pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu")
    W, b = cases.mlp_params(95, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
    pol = pol.to("cuda")
    qf = mirl_b200.B200EnsembleQValue(FakeEnv(), ensemble_size=E, hidden_layers=[256, 256], activation="relu")
    cW, cb = cases.mlp_params(96, E, [cases.OBS + cases.ACT, 256, 256, 1])
    qf.load_state_dict({k: T(v) for k, v in cases.mlp_state_dict(cW, cb).items()})
    qf = qf.to("cuda")
    before = qf.module.flat.data.clone()
    tr = mirl_b200.B200ActorTrainer(pol, qf, policy_lr=3e-4, target_entropy=-1.0,
                                    current_v_pi_kwargs={"sample_number": None, "mode": mode})
    steps = len(g["policy_loss"])
    for s_ in range(steps):
        o, _ = cases.rollout_inputs(970 + s_, 256)
        info = tr.update(T(o, "cuda"), noise=T(cases.actor_noise(s_), "cuda"))
        for key, name in (("policy/loss", "policy_loss"), ("policy/q_pi", "q_pi"), ("policy/entropy", "entropy"),
                          ("alpha/value", "alpha_value"), ("alpha/loss", "alpha_loss")):
            assert abs(info[key] - g[name][s_]) <= 1e-3 * abs(g[name][s_]) + 1e-5, (s_, key, info[key], g[name][s_])
        assert abs(float(tr.log_alpha.item()) - g["log_alpha"][s_]) <= 2e-6, (s_, float(tr.log_alpha.item()))
        if s_ + 1 in (1, steps):
            sd = pol.state_dict()
            msd = dict(zip([k for k, _ in pol._reference_items()],
                           [v for _, v in _policy_views(pol, tr.exp_avg)]))
            for i, name in enumerate(g["names"]):
                name = str(name)
                got, ref = proj(sd[name], 8000 + i), g["param_proj_%d" % (s_ + 1)][i]
                assert abs(got[0] - ref[0]) <= 3e-4 * (s_ + 1) * 0.05 * np.sqrt(sd[name].numel()) + 1e-4 * ref[0], (name, got, ref)
                gm, rm = proj(msd[name], 8100 + i), g["m_proj_%d" % (s_ + 1)][i]
                assert abs(gm[0] - rm[0]) <= 3e-3 * rm[0] + 2e-5, ("exp_avg", name, gm, rm)
    assert torch.equal(qf.module.flat.data, before)               # the critics are read, not updated


@pytest.mark.parametrize("kind", ["tqc", "subset", "fixed_alpha"])
def test_actor_update_variants_match_oracle(kind, monkeypatch):
    """The other reductions of the actor update against critic_oracle.actor_update (pinned to the reference by the
    actor_update fixtures): TQC critics (3x512, 25 quantiles: q_pi = mean of all 125 atoms), a batch-wise drawn
    subset (3 of 10 members, max), and a fixed entropy coefficient (use_automatic_entropy_tuning False)."""
    import mirl_b200
    from oracle import critic_oracle as co
    from tests.helpers_gpu import make_critic
    pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu")
    W, b = cases.mlp_params(97, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
    pol = pol.to("cuda")
    qf, (cW, cb) = make_critic("tqc" if kind == "tqc" else "redq", 98)
    kw = {"tqc": {"number_drop": 0}, "subset": {"sample_number": 3, "batchwise_sample": True, "mode": "max"},
          "fixed_alpha": {"sample_number": None, "mode": "min"}}[kind]
    auto = kind != "fixed_alpha"
    tr = mirl_b200.B200ActorTrainer(pol, qf, policy_lr=1e-3, target_entropy=-3.0, current_v_pi_kwargs=kw,
                                    use_automatic_entropy_tuning=auto, alpha_if_not_automatic=0.05)
    pW, pb = [T(a.copy()) for a in W], [T(a.copy()) for a in b]
    popt = {"step": 0, "m": [torch.zeros_like(t) for t in pW + pb], "v": [torch.zeros_like(t) for t in pW + pb]}
    log_alpha = torch.zeros(1) if auto else None
    aopt = {"step": 0, "m": [torch.zeros(1)], "v": [torch.zeros(1)]}
    rs = np.random.RandomState(12)
    for s_ in range(3):
        o, _ = cases.rollout_inputs(990 + s_, 200)
        eps = rs.standard_normal((200, cases.ACT)).astype(np.float32)
        members = None
        if kind == "subset":
            np.random.seed(40 + s_)
            members = np.random.choice(10, 3, replace=False)           # the draw the trainer makes (ensemble_value.py:85)
            np.random.seed(40 + s_)
        want = co.actor_update(pW, pb, popt, log_alpha, aopt, [T(a) for a in cW], [T(a) for a in cb], T(o), T(eps),
                               lr=1e-3, target_entropy=-3.0, mode=kw.get("mode", "mean"), member_idx=members,
                               alpha_fixed=0.05)
        info = tr.update(T(o, "cuda"), noise=T(eps, "cuda"))
        for key, v in want.items():
            assert abs(info[key] - v) <= 1e-3 * abs(v) + 1e-5, (s_, key, info[key], v)
    sd = pol.state_dict()
    for l in range(3):
        # Adam moves every weight by about lr per step whatever the gradient's size: compare on that scale
        assert (sd["module.net.%d.weight" % (2 * l)].cpu() - pW[l][0]).abs().max() <= 0.15 * 3 * 1e-3
        assert (sd["module.net.%d.bias" % (2 * l)].cpu() - pb[l][0]).abs().max() <= 0.15 * 3 * 1e-3


@pytest.mark.parametrize("tag", ["redq", "tqc"])
def test_sac_updater_matches_reference_agent(tag):
    """B200SACUpdater against the UNMODIFIED reference agents' train_from_torch_batch, replaying their RNG streams
    (np.random.seed(31) for the member draws, torch.manual_seed(32) for the policy draws).
    redq: SACAgent with configs/redq.json's settings, 24 updates -- TD target (policy sample + min over a random pair
    of target critics), critic update (MSE + Adam + soft target update), and at updates 0 and 20 the actor and alpha
    update.  tqc: TQCAgent with configs/tqc.json (truncated pooled atoms as the target, quantile-Huber critic loss on
    3x512 critics, actor through the mean of all atoms at every update), 6 updates."""
    import mirl_b200
    from tests.helpers_gpu import make_critic
    g = golden("sac_update_%s.npz" % tag)
    pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu")
    W, b = cases.mlp_params(95, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
    pol = pol.to("cuda")
    qf, _ = make_critic(tag, 71)
    qt, _ = make_critic(tag, 171)
    noise_fn = lambda shape: torch.randn(*shape).cuda()      # noqa: E731  (the reference's CPU generator, in call order)
    if tag == "redq":
        up = mirl_b200.B200SACUpdater(pol, qf, qt, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                                      policy_update_freq=20, target_entropy=-1, alpha_if_not_automatic=0,
                                      current_v_pi_kwargs={"sample_number": None, "mode": "mean"},
                                      next_v_pi_kwargs={"batchwise_sample": True}, noise_fn=noise_fn)
        freq = 20
    else:
        up = mirl_b200.B200SACUpdater(pol, qf, qt, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                                      alpha_if_not_automatic=0, noise_fn=noise_fn)          # TQCAgent's defaults
        freq = 1
    np.random.seed(31)
    torch.manual_seed(32)
    na, steps = 0, len(g["qf_loss"])
    for s_ in range(steps):
        x = cases.sac_batch(1200 + s_, 256)
        info = up.train_from_torch_batch({k: T(v, "cuda") for k, v in x.items()})
        assert abs(info["qf/loss"] - g["qf_loss"][s_]) <= 2e-3 * g["qf_loss"][s_], (s_, info["qf/loss"], g["qf_loss"][s_])
        if s_ % freq == 0:
            for k in ("policy/loss", "policy/q_pi", "policy/entropy", "alpha/value", "alpha/loss"):
                ref = g[k.replace("/", "_")][na]
                assert abs(info[k] - ref) <= 2e-3 * abs(ref) + 1e-4, (s_, k, info[k], ref)
            na += 1
    assert abs(float(up.actor.log_alpha.item()) - float(g["log_alpha"])) <= 3e-6
    for what, mod, key, base in (("qf", qf, "qf_proj", 9000), ("qf_target", qt, "qt_proj", 9100)):
        for i, p in enumerate(mod.parameters()):
            got, ref = proj(p, base + i), g[key][i]
            assert abs(got[0] - ref[0]) <= 3e-4 * steps * 0.05 * np.sqrt(p.numel()) + 1e-4 * ref[0], (what, i, got, ref)
    sd = pol.state_dict()
    for i, (name, _) in enumerate(pol._reference_items()):
        got, ref = proj(sd[name], 9200 + i), g["pol_proj"][i]
        assert abs(got[0] - ref[0]) <= 3e-4 * na * 0.05 * np.sqrt(sd[name].numel()) + 1e-4 * ref[0], (name, got, ref)


def _policy_views(pol, flat):
    """Per-tensor views of a blob laid out like the policy's parameters, in the reference's key order."""
    net = pol.net
    for l in range(net.num_fc):
        yield "w", net.weight(l, flat=flat)[0]
        yield "b", net.bias(l, flat=flat)[0].view(1, -1)


@pytest.mark.parametrize("B,act_fn,hidden", [(4096, "relu", [256, 256]), (10_037, "relu", [256, 256]),
                                           (6000, "swish", [64, 200, 32])])
def test_policy_one_launch_forward_matches_oracle(B, act_fn, hidden):
    """Rollout batch sizes take the one-launch forward (mlp_rows.cu: 64-row blocks, activations resident in shared
    memory): action, log-prob, mean and std against critic_oracle.policy_action -- a multiple of the block, a ragged
    row count, and other widths / activation / depth."""
    import mirl_b200
    from oracle import critic_oracle as co
    pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=hidden, activation=act_fn)
    pW, pb = cases.mlp_params(93, 1, [cases.OBS] + hidden + [2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(pW, pb).items()})
    pol = pol.to("cuda")
    o, _ = cases.rollout_inputs(94, B)
    eps = np.random.RandomState(95).standard_normal((B, cases.ACT)).astype(np.float32)
    a, info = pol.action(T(o, "cuda"), return_log_prob=True, return_mean_std=True, noise=T(eps, "cuda"))
    oa, olp, om, osd = co.policy_action([T(x) for x in pW], [T(x) for x in pb], T(o), T(eps), activation=act_fn)
    assert_close(a, oa, what="action")
    assert_close(info["mean"], om, what="mean")
    assert_close(info["std"], osd, what="std")
    assert_close(info["log_prob"], olp, what="log_prob")
    with pol.deterministic_(True):
        a, _ = pol.action(T(o, "cuda"))
    assert_close(a, torch.tanh(om), what="deterministic action")


# ---- N1: policy action -------------------------------------------------------------------------------
def test_gaussian_policy_action_matches_reference():
    """B200GaussianPolicy.action against the reference GaussianPolicy (17 -> 256 -> 256 -> 12 relu, tanh-squashed):
    reparameterised / plain samples from the recorded draw, log-prob, mean, std, deterministic action; state-dict
    keys are the reference's."""
    import mirl_b200
    g = golden("policy.npz")
    pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu")
    W, b = cases.mlp_params(91, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
    pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(W, b).items()})
    pol = pol.to("cuda")
    assert sorted(pol.state_dict().keys()) == [str(k) for k in g["keys"]]
    o, _ = cases.rollout_inputs(92, 300)
    for tag, rep in (("rsample", True), ("sample", False)):
        a, info = pol.action(T(o, "cuda"), return_log_prob=True, reparameterize=rep, return_mean_std=True,
                             noise=T(g["eps"], "cuda"))
        assert_close(a, g["action_" + tag], what="action " + tag)
        assert_close(info["log_prob"], g["log_prob_" + tag], what="log_prob " + tag)
        assert_close(info["mean"], g["mean"], what="mean")
        assert_close(info["std"], g["std"], what="std")
    with pol.deterministic_(True):
        a, _ = pol.action(T(o, "cuda"))
    assert_close(a, g["action_deterministic"], what="deterministic action")
    # without noise= the draw comes from the device generator: same distribution, right shapes, in (-1, 1)
    a, info = pol.action(T(o, "cuda"), return_log_prob=True)
    assert a.shape == (300, cases.ACT) and info["log_prob"].shape == (300, 1) and float(a.abs().max()) < 1.0
    # ragged / large batches through the tile kernels
    big = torch.randn(70001, cases.OBS, device="cuda")
    with pol.deterministic_(True):
        a1, _ = pol.action(big)
        a2, _ = pol.action(big[:4099])
    assert torch.allclose(a1[:4099], a2, atol=2e-5)
