"""GPU parity tests: the CUDA path (through the C ABI) against the committed reference fixtures
and the CPU oracle on the same seeded inputs.  Tolerance of record: tests/util.py RTOL = 1e-3
(|err| <= 1e-3*|ref| + 1e-3*mean|ref|); done masks and member indices bit-exact (done flips are
tolerated only inside a 1e-3 guard band around a threshold and are counted)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import critic_oracle as co
from oracle import pe_oracle as po
from tests.golden import cases
from tests.helpers_gpu import done_guard_mismatches, make_critic, make_pe_model
from tests.util import T, assert_close, golden, norm_rel, proj, to_torch

pytestmark = pytest.mark.gpu
LOSS_VEC7 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4]
LOSS_VEC10 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4, 0.05, 0.8, 0.6]
KERNELS = ["simt", "tcgen05"]


def _kernel_or_skip(m, kernel):
    if kernel == "tcgen05" and not m._tc_supported():
        pytest.skip("tcgen05 kernel not built for this shape")


@pytest.mark.parametrize("kernel", KERNELS)
def test_forward_matches_reference_fixture(kernel):
    g = golden("pe_forward_mbpo.npz")
    m, _ = make_pe_model(cases.MBPO, 11, kernel=kernel)
    _kernel_or_skip(m, kernel)
    o, a = cases.rollout_inputs(12, 200)
    mean, ls = m.predict_distribution(T(o, "cuda"), T(a, "cuda"))
    assert mean.shape == ls.shape == (7, 200, 18)
    assert_close(mean, g["mean"], what="mean")
    assert_close(ls, g["logstd"], what="logstd")
    # margin against the float64 reference: SIMT fp32 is the reference's own arithmetic class,
    # the tensor-core path is the bf16x3 split (~2^-16 per product)
    assert norm_rel(mean, g["mean64"]) < (1e-5 if kernel == "simt" else 2e-5)


@pytest.mark.parametrize("B", [1, 127, 129, 4099])
def test_forward_tc_ragged_batches_match_simt(B):
    """All-member forward on the tensor cores (rollout kernel in forward mode, tiles walk every
    member's rows in order) against the SIMT forward, shared and per-member inputs."""
    ms, _ = make_pe_model(cases.MBPO, 11, kernel="simt")
    mt, _ = make_pe_model(cases.MBPO, 11, kernel="tcgen05")
    _kernel_or_skip(mt, "tcgen05")
    o, a = cases.rollout_inputs(900 + B, B)
    for oo, aa in ((T(o, "cuda"), T(a, "cuda")),
                   (T(o, "cuda")[None].repeat(7, 1, 1) * torch.arange(1, 8, device="cuda")[:, None, None] * 0.3,
                    T(a, "cuda")[None].repeat(7, 1, 1))):
        m1, l1 = ms.predict_distribution(oo, aa)
        m2, l2 = mt.predict_distribution(oo, aa)
        assert_close(m2, m1, what="mean")
        assert_close(l2, l1, what="logstd")


@pytest.mark.parametrize("kernel", KERNELS)
def test_forward_per_member_inputs_match_oracle(kernel):
    m, p = make_pe_model(cases.MBPO, 11, kernel=kernel)
    _kernel_or_skip(m, kernel)
    bt = cases.train_batch(41, 7, 129)
    pt = to_torch(p)
    x = torch.cat([po.normalize(T(bt["observations"]), pt["obs_mean"], pt["obs_std"]),
                   po.normalize(T(bt["actions"]), pt["act_mean"], pt["act_std"])], -1)
    ref_mean, ref_ls = po.mean_logstd(po.mlp_forward(pt["W"], pt["b"], x), pt["max_logstd"], pt["min_logstd"])
    mean, ls = m.predict_distribution(T(bt["observations"], "cuda"), T(bt["actions"], "cuda"))
    assert_close(mean, ref_mean, what="mean")
    assert_close(ls, ref_ls, what="logstd")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("env", po.DONE_KINDS)
def test_rollout_step_matches_reference_fixture(env, kernel):
    g = golden("pe_step_%s.npz" % env)
    seed = 21 + po.DONE_KINDS.index(env)
    m, _ = make_pe_model(cases.MBPO, 11, env, kernel=kernel)
    _kernel_or_skip(m, kernel)
    m.remember_loss(LOSS_VEC7)
    assert list(m.elite_indices) == list(g["elite_indices"])
    o, a = cases.rollout_inputs(seed, 1000, env=env)
    # replay of the reference's RNG consumption: same seeds -> same member indices
    np.random.seed(seed)
    idx = np.random.choice(m.elite_indices, 1000)
    assert np.array_equal(idx, g["idx"])
    no, r, d, info = m.step(T(o, "cuda"), T(a, "cuda"), indices=idx, noise=T(g["eps"], "cuda"))
    assert info == {}
    assert_close(no, g["next_obs"], what="next_obs")
    assert_close(r, g["reward"], what="reward")
    outside, inside = done_guard_mismatches(env, g["next_obs"], g["done"], d.cpu().numpy())
    assert outside == 0, "%d done flags differ outside the guard band (%d inside)" % (outside, inside)
    assert inside <= 2
    if env == "cheetah":
        assert float(d.abs().sum()) == 0.0


@pytest.mark.parametrize("kernel", KERNELS)
def test_rollout_step_without_elites_draws_like_reference(kernel):
    g = golden("pe_step_noelite.npz")
    m, _ = make_pe_model(cases.MBPO, 11, kernel=kernel)
    _kernel_or_skip(m, kernel)
    o, a = cases.rollout_inputs(27, 64)
    np.random.seed(27)                      # step() draws np.random.choice(arange(E), B) itself
    no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), noise=T(g["eps"], "cuda"))
    assert_close(no, g["next_obs"], what="next_obs")
    assert_close(r, g["reward"], what="reward")


@pytest.mark.parametrize("kernel", KERNELS)
def test_return_distribution_and_mopo_penalty(kernel):
    g = golden("pe_step_dist_mopo.npz")
    m, _ = make_pe_model(cases.MOPO, 31, "hopper", kernel=kernel)
    _kernel_or_skip(m, kernel)
    m.remember_loss(LOSS_VEC10)
    assert list(m.elite_indices) == list(g["elite_indices"])
    o, a = cases.rollout_inputs(32, 128, env="hopper")
    for kind in ("total_var", "max_var"):
        no, r, d, info = m.step(T(o, "cuda"), T(a, "cuda"), return_distribution=True, indices=g["idx"],
                                noise=T(g["eps"], "cuda"), penalty_type=kind)
        for k in ("ensemble_delta_mean", "ensemble_delta_std", "ensemble_next_obs_mean",
                  "ensemble_next_obs_std", "ensemble_reward_mean", "ensemble_reward_std"):
            assert_close(info[k], g[k], what=k)
        assert_close(no, g["next_obs"], what="next_obs")
        assert_close(r, g["reward"], what="reward")
        assert_close(info["penalized_reward"], g["penalized_reward_" + kind], what="penalized " + kind)
        outside, _ = done_guard_mismatches("hopper", g["next_obs"], g["done"], d.cpu().numpy())
        assert outside == 0
    assert list(info.keys())[:6] == ["ensemble_delta_mean", "ensemble_next_obs_mean", "ensemble_delta_std",
                                     "ensemble_next_obs_std", "ensemble_reward_mean", "ensemble_reward_std"]


def _batch(seed, E, B, shared=False, device="cuda"):
    return {k: T(v, device) for k, v in cases.train_batch(seed, E, B, shared=shared).items()}


def test_compute_loss_matches_reference_fixture():
    g = golden("pe_loss_mbpo.npz")
    m, _ = make_pe_model(cases.MBPO, 11)
    bt = _batch(41, 7, 256)
    for tag, ignore in (("keep", False), ("ignore", True)):
        loss, el = m.compute_loss(bt["observations"], bt["actions"], bt["deltas"], bt["rewards"],
                                  bt["terminals"], ignore)
        assert_close(loss, g["loss_" + tag], what="loss " + tag)
        assert_close(el, g["ensemble_loss_" + tag], what="ensemble_loss " + tag)
    bs = _batch(42, 7, 500, shared=True)
    loss, el = m.compute_loss(bs["observations"], bs["actions"], bs["deltas"], bs["rewards"], bs["terminals"])
    assert_close(el, g["ensemble_loss_shared"], what="shared ensemble_loss")
    assert_close(loss, g["loss_shared"], what="shared loss")


def _grads(m, bt, ignore):
    from mirl_b200 import _lib
    L = _lib.lib()
    cm = m._c_model(None)
    cfg = m.train_cfg(ignore)
    B = bt["observations"].shape[1]
    ws = torch.empty(L.mb_pe_train_workspace_bytes(ctypes.byref(cm), B), dtype=torch.uint8, device="cuda")
    grads = torch.zeros_like(m.module.flat.data)
    el = torch.zeros(m.ensemble_size, device="cuda")
    terms = torch.zeros(3, device="cuda")
    _lib.check(L.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(cfg), _lib.ptr(bt["observations"]),
                                  _lib.ptr(bt["actions"]), _lib.ptr(bt["deltas"]), _lib.ptr(bt["rewards"]),
                                  _lib.ptr(bt["terminals"]), B, _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms),
                                  _lib.ptr(ws), ws.numel(), _lib.stream()))
    torch.cuda.synchronize()
    return grads, el, terms


def test_gradients_match_reference_autograd():
    g = golden("pe_loss_small.npz")
    m, _ = make_pe_model(cases.SMALL, 51)
    grads, el, terms = _grads(m, _batch(52, 3, 64), False)
    assert_close(terms.sum(), g["loss"], what="loss")
    ref_sd = m.module.reference_state_dict("module.", flat=grads)
    for k in g.files:
        if k.startswith("grad/"):
            assert norm_rel(ref_sd[k[5:]], g[k]) < 1e-3, k
    g = golden("pe_loss_mbpo.npz")
    m, _ = make_pe_model(cases.MBPO, 11)
    for tag, ignore in (("keep", False), ("ignore", True)):
        grads, el, terms = _grads(m, _batch(41, 7, 256), ignore)
        assert_close(el, g["ensemble_loss_" + tag], what="ensemble_loss")
        assert_close(terms.sum(), g["loss_" + tag], what="loss")
        sd = m.module.reference_state_dict("module.", flat=grads)
        for i, name in enumerate(g["names"]):
            got, ref = proj(sd[str(name)], 1000 + i), g["grad_proj_" + tag][i]
            assert abs(got[0] - ref[0]) <= 1e-3 * ref[0] + 1e-9, (name, got, ref)     # norm
            assert abs(got[1] - ref[1]) <= 1e-3 * ref[0] + 1e-9, (name, got, ref)     # projection


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("tag,cfg,seed,B", [("small", cases.SMALL, 51, 64), ("mbpo", cases.MBPO, 11, 256)])
def test_fused_train_steps_match_reference_adam(tag, cfg, seed, B, graph):
    """Ten fused steps against the reference's Adam run; stream launches (programmatic dependent
    launch, host-side step count) and CUDA-graph replay (device-side step count)."""
    import mirl_b200
    from tests.helpers_gpu import FakeEnv
    g = golden("pe_train_%s.npz" % tag)
    m, _ = make_pe_model(cfg, seed)
    tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, lr=1e-3, tb_log_freq=-1, silent=True,
                                      ignore_terminal_state=False, use_cuda_graph=graph)
    for s in range(10):
        diag = tr.train_from_torch_batch(_batch(600 + s, cfg["E"], B))
        assert_close(np.array(diag["train_loss"]), g["loss"][s], what="loss step %d" % s)
        el = np.array([diag["mt/net_%d/train_loss" % i] for i in range(cfg["E"])])
        assert_close(el, g["ensemble_loss"][s], what="ensemble_loss step %d" % s)
        if s + 1 in (1, 10):
            sd = m.state_dict()
            msd = m.module.reference_state_dict("module.", flat=tr.exp_avg)
            vsd = m.module.reference_state_dict("module.", flat=tr.exp_avg_sq)
            for i, name in enumerate(g["names"]):
                name = str(name)
                got, ref = proj(sd[name], 2000 + i), g["param_proj_%d" % (s + 1)][i]
                # Adam's early steps are sign-like: absolute tolerance ~ lr * sqrt(numel)
                tol = 1e-3 * (s + 1) * 0.05 * np.sqrt(sd[name].numel()) + 1e-4 * ref[0]
                assert abs(got[0] - ref[0]) <= tol, (name, got, ref)
                gm, rm = proj(msd[name], 3000 + i), g["m_proj_%d" % (s + 1)][i]
                assert abs(gm[0] - rm[0]) <= 2e-3 * rm[0] + 1e-9, ("exp_avg", name, gm, rm)
                gv, rv = proj(vsd[name], 4000 + i), g["v_proj_%d" % (s + 1)][i]
                assert abs(gv[0] - rv[0]) <= 4e-3 * rv[0] + 1e-12, ("exp_avg_sq", name, gv, rv)
            if tag == "small":
                for k in g.files:
                    if k.startswith("param_%d/" % (s + 1)):
                        err = np.abs(sd[k.split("/", 1)[1]].cpu().numpy() - g[k]).max()
                        assert err < 5e-4, (k, err)


def test_redq_target_matches_reference_fixture():
    g = golden("redq.npz")
    q, _ = make_critic("redq", 71)
    x = cases.critic_inputs(72, 256)
    o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
    _, info = q.value(o, a, only_return_ensemble=True)
    assert_close(info["ensemble_value"], g["ensemble"], what="ensemble")
    np.random.seed(73)
    v, _ = q.value(o, a, sample_number=2, batchwise_sample=True, mode="min")
    assert_close(v, g["bw_min"], what="batch-wise min")
    np.random.seed(74)
    v, _ = q.value(o, a, sample_number=2, batchwise_sample=False, mode="min")
    assert_close(v, g["row_min"], what="per-row min")
    np.random.seed(75)
    v, _ = q.value(o, a, sample_number=3, batchwise_sample=False, mode="max")
    assert_close(v, g["row3_max"], what="per-row max")
    assert_close(q.value(o, a, sample_number=None, mode="mean")[0], g["all_mean"], what="mean")
    assert_close(q.value(o, a, sample_number=10, mode="min")[0], g["all_min"], what="min")
    np.random.seed(73)
    y = q.q_target(o, a, T(x["rewards"], "cuda"), T(x["terminals"], "cuda"), T(x["log_prob"], "cuda"),
                   alpha=0.2, discount=0.99, sample_number=2, batchwise_sample=True)
    assert_close(y, g["td_target"], what="td target")


@pytest.mark.parametrize("B", [1, 37, 129, 300, 1000])
def test_critic_ragged_batches_match_oracle(B):
    """REDQ (2x256) and TQC (3x512 x 25 atoms) forward at batch sizes that are not multiples of
    the 128-row UMMA tile: hidden/output layers run on the tcgen05 GEMM (partial row tiles, N
    tiles of 128/128/16/32 columns, reductions of 4 and 8 stages)."""
    x = cases.critic_inputs(700 + B, B)
    o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
    q, (W, b) = make_critic("redq", 71)
    _, info = q.value(o, a, only_return_ensemble=True)
    ref = co.ensemble_q([T(w) for w in W], [T(v) for v in b], T(x["obs"]), T(x["act"]))
    assert_close(info["ensemble_value"], ref, what="redq ensemble")
    t, (W, b) = make_critic("tqc", 81)
    v, info = t.value(o, a, return_atoms=True, percentage_drop=8)
    rv, ratoms = co.tqc_atoms([T(w) for w in W], [T(v_) for v_ in b], T(x["obs"]), T(x["act"]), percentage_drop=8)
    assert_close(v, rv, what="tqc value")
    assert_close(info["value_atoms"], ratoms, what="tqc atoms")


@pytest.mark.parametrize("B", [5, 100, 257])
def test_train_gradients_ragged_batches_match_oracle(B):
    """Analytic gradients through the tcgen05 forward / dX / dW GEMMs at batch sizes that leave
    partial 128-row tiles and partial 64-element reduction stages in the dW GEMM."""
    m, p = make_pe_model(cases.MBPO, 11)
    bt = _batch(43 + B, 7, B)
    grads, el, terms = _grads(m, bt, False)
    pt = to_torch(p)
    cb = {k: v.cpu() for k, v in bt.items()}
    ref = po.model_loss_grads(pt, cb["observations"], cb["actions"], cb["deltas"], cb["rewards"], cb["terminals"],
                              ignore_terminal_state=False, weight_decay=cases.WEIGHT_DECAY)
    assert_close(terms.sum(), ref["loss"], what="loss")
    assert_close(el, ref["ensemble_loss"], what="ensemble loss")
    mod = m.module
    for l in range(5):
        assert norm_rel(mod.weight(l, flat=grads), ref["W"][l]) < 1e-3, "dW%d" % l
        assert norm_rel(mod.bias(l, flat=grads).reshape(ref["b"][l].shape), ref["b"][l]) < 1e-3, "db%d" % l
    assert norm_rel(mod.logstd_bound(0, flat=grads).reshape(ref["max_logstd"].shape), ref["max_logstd"]) < 1e-3
    assert norm_rel(mod.logstd_bound(1, flat=grads).reshape(ref["min_logstd"].shape), ref["min_logstd"]) < 1e-3


def test_tqc_target_matches_reference_fixture():
    g = golden("tqc.npz")
    q, _ = make_critic("tqc", 81)
    x = cases.critic_inputs(82, 256)
    o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
    _, info = q.value(o, a, only_return_ensemble=True)
    assert_close(info["ensemble_atoms"], g["ensemble"], what="ensemble atoms")
    v, info = q.value(o, a, percentage_drop=8, return_atoms=True)
    assert info["value_atoms"].shape == (256, 115)
    assert_close(info["value_atoms"], g["atoms_drop8pct"], what="kept atoms")
    assert_close(v, g["value_drop8pct"], what="value")
    at = info["value_atoms"]
    assert bool((at[:, 1:] >= at[:, :-1]).all())                 # ascending
    v, info = q.value(o, a, number_drop=0, return_atoms=True)
    assert_close(info["value_atoms"], g["atoms_drop0"], what="atoms drop 0")
    y = q.atoms_target(o, a, T(x["rewards"], "cuda"), T(x["terminals"], "cuda"), T(x["log_prob"], "cuda"),
                       alpha=0.2, discount=0.99, percentage_drop=8)
    assert_close(y, g["atoms_target"], what="atoms target")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("B", [1, 63, 64, 65, 127, 128, 129, 257, 4099])
def test_rollout_ragged_batches_match_oracle(B, kernel):
    m, p = make_pe_model(cases.MBPO, 11, "hopper", kernel=kernel)
    _kernel_or_skip(m, kernel)
    m.remember_loss(LOSS_VEC7)
    o, a = cases.rollout_inputs(100 + B, B, env="hopper")
    rs = np.random.RandomState(B)
    idx = rs.choice(m.elite_indices, B)
    eps = rs.standard_normal((B, 18)).astype(np.float32)
    no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), indices=idx, noise=T(eps, "cuda"))
    rno, rr, rd = po.rollout_step(to_torch(p), T(o), T(a), torch.from_numpy(idx), T(eps), "hopper")
    assert_close(no, rno, what="next_obs")
    assert_close(r, rr, what="reward")
    outside, _ = done_guard_mismatches("hopper", rno.numpy(), rd.numpy(), d.cpu().numpy())
    assert outside == 0


def _sorted_rows(t):
    a = t.detach().cpu().numpy()
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("env,B", [("hopper", 1), ("hopper", 129), ("cheetah", 5000), ("walker2d", 4099)])
def test_step_rows_equals_step_plus_pool_assembly(env, B, kernel):
    """The fused pool-row output is bit-identical to step() followed by the collector's
    concatenation [o | a | o' | r | d] (mbpo_collector.py:69-83) -- in two destinations at a row
    offset, with and without row padding; again when the next horizon step reads obs as a strided
    view of those rows; and as the same multiset of rows in bucket order (ordered=False)."""
    m, _ = make_pe_model(cases.MBPO, 11, env, kernel=kernel)
    _kernel_or_skip(m, kernel)
    m.remember_loss(LOSS_VEC7)
    O, A = m.obs_size, m.action_size
    W, S = m.row_width(), m.row_stride()
    assert (W, S) == (42, 44)
    o, a = cases.rollout_inputs(300 + B, B, env=env)
    o, a = T(o, "cuda"), T(a, "cuda")
    rs = np.random.RandomState(B)
    idx = rs.choice(m.elite_indices, B)
    eps = T(rs.standard_normal((B, O + 1)).astype(np.float32), "cuda")
    no, r, d, _ = m.step(o, a, indices=idx, noise=eps)
    want = torch.cat([o, a, no, r, d], 1)
    off = 8
    for stride in (W, S):
        dst = [torch.full((B + 11, stride), -7.0, device="cuda") for _ in range(2)]
        vno, vr, vd = m.step_rows(o, a, dst, row_offset=off, indices=idx, noise=eps)
        for t in dst:
            assert torch.equal(t[off:off + B, :W], want)
            assert float(t[off:off + B, W:].abs().sum()) == 0
            assert float((t[:off] + 7).abs().sum()) == 0 and float((t[off + B:] + 7).abs().sum()) == 0
        assert torch.equal(vno, no) and torch.equal(vr, r) and torch.equal(vd, d)
    # chained step: obs is the next_obs column block of the rows just written (row stride S)
    a2 = T(cases.rollout_inputs(301 + B, B, env=env)[1], "cuda")
    no2, r2, d2, _ = m.step(no, a2, indices=idx, noise=eps)
    dst2 = torch.empty(B, W, device="cuda")
    m.step_rows(vno, a2, dst2, indices=idx, noise=eps)
    assert torch.equal(dst2, torch.cat([no, a2, no2, r2, d2], 1))
    # bucket order: same rows, grouped by member, every destination identical
    dst3 = [torch.full((B + 11, S), -7.0, device="cuda") for _ in range(2)]
    m.step_rows(o, a, dst3, row_offset=off, indices=idx, noise=eps, ordered=False)
    assert torch.equal(dst3[0], dst3[1])
    got = dst3[0][off:off + B]
    assert float(got[:, W:].abs().sum()) == 0
    assert float((dst3[0][:off] + 7).abs().sum()) == 0 and float((dst3[0][off + B:] + 7).abs().sum()) == 0
    assert np.array_equal(_sorted_rows(got[:, :W]), _sorted_rows(want))
    # grouped by member: the member of each output row (recovered through its obs) is non-decreasing
    # in the order the members appear in the elite list's numeric order
    key = {tuple(np.round(x, 6)): int(i) for x, i in zip(o.cpu().numpy(), idx)}
    mem = np.array([key[tuple(np.round(x, 6))] for x in got[:, :O].cpu().numpy()])
    assert np.all(np.diff(mem) >= 0)


@pytest.mark.parametrize("seed,elites,B", [(0, [3, 1, 5, 0, 6], 1000), (1, [2, 0, 1, 4, 3, 6, 5], 100003),
                                           (5, [0, 1], 7), (9, [4], 10), (3, list(range(15)), 50000),
                                           (7, [6, 2, 4, 1, 0], 1_000_000)])
def test_device_member_draw_is_numpy_choice_bit_exact(seed, elites, B):
    """draw_members == np.random.choice(elite, B) (pe_model.py:169) bit for bit, and the global
    NumPy generator ends in the same state (next host draws agree, cached Gaussian included)."""
    m, _ = make_pe_model(cases.SMALL, 11)            # only the elite list matters here
    m.elite_indices = list(elites)
    np.random.seed(seed)
    np.random.standard_normal(3)                     # leaves a cached Gaussian in the state
    st = np.random.get_state()
    want = np.random.choice(elites, B)
    want_next = np.random.standard_normal(5), np.random.randint(0, 1000, 5)
    np.random.set_state(st)
    got = m.draw_members(B)
    got_next = np.random.standard_normal(5), np.random.randint(0, 1000, 5)
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), want.astype(np.uint8))
    assert np.array_equal(want_next[0], got_next[0]) and np.array_equal(want_next[1], got_next[1])


def test_speculative_member_draws_follow_the_numpy_stream():
    """Repeated draws of one shape are generated one call ahead (pe_model.draw_members); every
    call must still return what np.random.choice returns at that point of the GLOBAL stream,
    also when other host code draws in between (the queued draw is then dropped), when the batch
    size changes, and when the stream is re-seeded."""
    m, _ = make_pe_model(cases.SMALL, 11)
    elites = [3, 1, 5, 0, 6]
    m.elite_indices = list(elites)
    plan = [40000, 40000, 40000, 40000, "host", 40000, 40000, 12345, 40000, 40000, "seed", 40000, 40000, 40000]
    np.random.seed(31)
    want = []
    for p in plan:
        if p == "host":
            want.append(np.random.standard_normal(3))
        elif p == "seed":
            np.random.seed(77)
        else:
            want.append(np.random.choice(elites, p).astype(np.uint8))
    want_tail = np.random.randint(0, 1 << 30, 4)
    np.random.seed(31)
    got, hits = [], 0
    for p in plan:
        if p == "host":
            got.append(np.random.standard_normal(3))
        elif p == "seed":
            np.random.seed(77)
        else:
            hits += len(m._draw_queue) > 0
            got.append(m.draw_members(p).cpu().numpy())
    assert hits >= 5                                  # the look-ahead actually ran
    assert all(np.array_equal(w, g) for w, g in zip(want, got))
    assert np.array_equal(want_tail, np.random.randint(0, 1 << 30, 4))


def test_step_draws_members_on_device_for_large_batches():
    m, _ = make_pe_model(cases.MBPO, 11)
    m.remember_loss(LOSS_VEC7)
    B = m.device_draw_min_rows + 5
    o, a = cases.rollout_inputs(77, B)
    o, a = T(o, "cuda"), T(a, "cuda")
    eps = torch.randn(B, 18, device="cuda")
    np.random.seed(123)
    idx = np.random.choice(m.elite_indices, B)
    want = m.step(o, a, indices=idx, noise=eps)
    np.random.seed(123)
    got = m.step(o, a, noise=eps)
    assert all(torch.equal(x, y) for x, y in zip(want[:3], got[:3]))


def test_rollout_empty_batch():
    m, _ = make_pe_model(cases.MBPO, 11)
    no, r, d, _ = m.step(torch.zeros(0, 17, device="cuda"), torch.zeros(0, 6, device="cuda"))
    assert no.shape == (0, 17) and r.shape == (0, 1) and d.shape == (0, 1)


@pytest.mark.parametrize("E,elites", [(2, 2), (5, 3), (10, 7), (20, 15)])
def test_ensemble_sizes_match_oracle(E, elites):
    cfg = dict(E=E, elites=elites, hidden=[200, 200, 200, 200], act="swish")
    m, p = make_pe_model(cfg, 200 + E, kernel="simt")
    m.remember_loss(list(np.random.RandomState(E).uniform(size=E)))
    B = 300
    o, a = cases.rollout_inputs(300 + E, B)
    rs = np.random.RandomState(E)
    idx = rs.choice(m.elite_indices, B)
    eps = rs.standard_normal((B, 18)).astype(np.float32)
    no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), indices=idx, noise=T(eps, "cuda"))
    rno, rr, _ = po.rollout_step(to_torch(p), T(o), T(a), torch.from_numpy(idx), T(eps))
    assert_close(no, rno, what="next_obs")
    assert_close(r, rr, what="reward")


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_properties(kernel):
    """BASELINE.json full size (1e6 states): size-independent properties instead of an oracle
    pass -- (a) rows are independent: a permuted batch gives the permuted result bit-for-bit;
    (b) zero noise reproduces obs + recover(mean of the drawn member) from the all-member
    forward kernel; (c) cheetah never terminates."""
    m, _ = make_pe_model(cases.MBPO, 11, kernel=kernel)
    _kernel_or_skip(m, kernel)
    m.remember_loss(LOSS_VEC7)
    B = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    o = torch.randn(B, 17, device="cuda", generator=g)
    a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
    eps = torch.randn(B, 18, device="cuda", generator=g)
    idx = torch.from_numpy(np.random.RandomState(5).choice(m.elite_indices, B)).cuda()
    no, r, d, _ = m.step(o, a, indices=idx, noise=eps)
    perm = torch.randperm(B, device="cuda", generator=g)
    no2, r2, d2, _ = m.step(o[perm], a[perm], indices=idx[perm], noise=eps[perm])
    assert torch.equal(no[perm], no2) and torch.equal(r[perm], r2) and torch.equal(d[perm], d2)
    assert float(d.sum()) == 0.0
    n = 20000
    no0, r0, _, _ = m.step(o[:n], a[:n], indices=idx[:n], noise=torch.zeros(n, 18, device="cuda"))
    mean, _ = m.predict_distribution(o[:n], a[:n])
    sel = mean[idx[:n].long(), torch.arange(n, device="cuda")]
    ref = o[:n] + m.delta_processor.recover(sel[:, :17])
    assert_close(no0, ref, what="zero-noise next_obs")
    assert_close(r0, sel[:, 17:], what="zero-noise reward")


def test_maximum_size_rollout_10m_rows():
    """BASELINE.json's largest rollout batch (1e7 states, horizon 15 chained in place): rows are
    independent, so every horizon step of a 50k-row slice taken alone must reproduce the same rows
    of the full run bit for bit (catches 32-bit index overflow in the bucketing, the tile plan and
    the pool-row addressing), and the pool rows must chain: obs of step h == next_obs of h-1."""
    m, _ = make_pe_model(cases.MBPO, 11, kernel="tcgen05")
    _kernel_or_skip(m, "tcgen05")
    m.remember_loss(LOSS_VEC7)
    B, H, O, A = 10_000_000, 15, 17, 6
    lo, hi = B - 50_000, B                       # the slice with the largest addresses
    g = torch.Generator(device="cuda").manual_seed(9)
    o = torch.randn(B, O, device="cuda", generator=g)
    pool = torch.empty(2 * B, m.row_stride(), device="cuda")
    small = torch.empty(2, hi - lo, m.row_stride(), device="cuda")
    os_ = o[lo:hi].clone()
    for h in range(H):
        a = torch.rand(B, A, device="cuda", generator=g) * 2 - 1
        eps = torch.randn(B, O + 1, device="cuda", generator=g)
        idx = torch.randint(0, 5, (B,), device="cuda", generator=g).to(torch.uint8)
        idx = torch.tensor(m.elite_indices, device="cuda", dtype=torch.uint8)[idx.long()]
        off = (h & 1) * B
        no, r, d = m.step_rows(o, a, pool, row_offset=off, indices=idx, noise=eps)
        rows = pool[off:off + B]
        assert torch.equal(rows[:, :O], o) and torch.equal(rows[:, O:O + A], a)
        sn, sr, sd = m.step_rows(os_, a[lo:hi], small[h & 1], indices=idx[lo:hi], noise=eps[lo:hi])
        assert torch.equal(no[lo:hi], sn) and torch.equal(r[lo:hi], sr) and torch.equal(d[lo:hi], sd)
        assert bool(torch.isfinite(no).all())
        o, os_ = no, sn                          # views into the pool rows: the next step reads them in place
    assert float(pool[:, 2 * O + A + 1].abs().sum()) == 0.0      # cheetah never terminates


def test_device_pool_matches_reference_simple_pool():
    """B200DevicePool against the reference SimplePool run (tests/golden/pool.npz): ring inserts that
    wrap, sampled batches from the same NumPy stream (bit-exact: pure data movement), get_data
    order after the wrap, resize."""
    import mirl_b200
    from tests.helpers_gpu import FakeEnv
    g = golden("pool.npz")
    pool = mirl_b200.B200DevicePool(FakeEnv("cheetah"), max_size=1000)
    np.random.seed(91)
    for i, n in enumerate(cases.POOL_ADDS):
        pool.add_samples(cases.pool_samples(910 + i, n))
        b = pool.random_batch(64)
        assert [pool.get_size(), pool._stop] == list(g["size_%d" % i])
        for k, v in b.items():
            assert np.array_equal(v, g["batch_%d_%s" % (i, k)]), (i, k)
    for k, v in pool.get_data().items():
        assert np.allclose(proj(torch.from_numpy(v), 7), g["data_" + k], rtol=1e-12), k
    pool.resize(1500)
    pool.add_samples(cases.pool_samples(990, 300))
    assert [pool.get_size(), pool._stop] == list(g["size_resized"])
    for k, v in pool.get_data().items():
        assert np.allclose(proj(torch.from_numpy(v), 8), g["data_resized_" + k], rtol=1e-12), k
    bt = pool.random_batch_torch(32, keys=["observations", "rewards"])
    assert all(v.is_cuda for v in bt.values())
    for k, v in bt.items():
        assert np.array_equal(v.cpu().numpy(), g["batch_resized_" + k]), k


def test_device_pool_filled_by_the_rollout_kernel():
    """reserve() + step_rows: the rollout kernel writes its transitions straight into the pool's
    ring; sampling then returns rows of the fused step (same multiset as step() + add_samples)."""
    import mirl_b200
    from tests.helpers_gpu import FakeEnv
    m, _ = make_pe_model(cases.MBPO, 11, "hopper")
    m.remember_loss(LOSS_VEC7)
    B = 1500
    pool = mirl_b200.B200DevicePool(FakeEnv("hopper"), max_size=4000, row_stride=m.row_stride())
    ref = mirl_b200.B200DevicePool(FakeEnv("hopper"), max_size=4000, row_stride=m.row_stride())
    o, a = cases.rollout_inputs(333, B, env="hopper")
    o, a = T(o, "cuda"), T(a, "cuda")
    rs = np.random.RandomState(3)
    for h in range(2):
        idx = rs.choice(m.elite_indices, B)
        eps = T(rs.standard_normal((B, 18)).astype(np.float32), "cuda")
        no, r, d, _ = m.step(o, a, indices=idx, noise=eps)
        ref.add_samples({"observations": o, "actions": a, "next_observations": no, "rewards": r, "terminals": d})
        start = pool.reserve(B)
        assert start == h * B
        m.step_rows(o, a, pool.rows, row_offset=start, indices=idx, noise=eps, ordered=False)
        o = no
    assert pool.reserve(B) is None and pool.get_size() == ref.get_size() == 2 * B
    for h in range(2):
        got = pool.rows[h * B:(h + 1) * B, :42]
        want = ref.rows[h * B:(h + 1) * B, :42]
        assert np.array_equal(_sorted_rows(got), _sorted_rows(want))


def test_collector_with_device_pool_equals_reference_shaped_path():
    """B200MBPOCollector over a B200DevicePool (rollout kernel inserts rows in place, no host copy)
    stores exactly the transitions the reference-shaped path (model.step + add_samples of host
    copies) stores, for hopper rollouts where trajectories terminate and are compacted away."""
    import mirl_b200
    from mirl_b200.collectors import B200MBPOCollector
    from tests.helpers_gpu import FakeEnv

    class Policy:
        def action(self, o):
            g = torch.Generator(device="cuda").manual_seed(int(o.shape[0]))
            return torch.rand(o.shape[0], 6, device="cuda", generator=g) * 2 - 1, {}

    class StartPool:
        def random_batch_torch(self, n, keys=None):
            o, _ = cases.rollout_inputs(444, n, env="hopper")
            return {"observations": T(o)}

    m, _ = make_pe_model(cases.MBPO, 11, "hopper")
    m.remember_loss(LOSS_VEC7)
    pools = []
    for kind in ("device", "reference-shaped"):
        im = mirl_b200.B200DevicePool(FakeEnv("hopper"), max_size=4 * 3000 * 20, row_stride=m.row_stride())
        c = B200MBPOCollector(m, Policy(), StartPool(), im, depth_schedule=4, number_sample=3000, bathch_size=1500)
        if kind != "device":
            c._fused_step = lambda o, a: None
        np.random.seed(5)
        torch.manual_seed(5)
        c.imagine(0)
        pools.append(im)
    a, b = pools
    assert a.get_size() == b.get_size() and 3000 < a.get_size() <= 4 * 3000
    assert torch.equal(a.rows[:a.get_size(), :42], b.rows[:b.get_size(), :42])
    assert float(a.rows[:a.get_size(), 41].sum()) > 0          # some trajectories did terminate


@pytest.mark.parametrize("obs,act,hidden,E", [(11, 3, [200, 200, 200, 200], 7),      # hopper: D = 12 (one head block)
                                               (27, 4, [128, 128, 128], 5),          # K0 = 31: widest layer-0 operand
                                               (3, 1, [64, 64], 2),                  # smallest widths the tile kernel takes
                                               (17, 6, [240, 240], 3),               # HP = 256: widest hidden layer
                                               (17, 6, [200, 200, 200, 200, 200, 200], 4)])
def test_tcgen05_paths_match_simt_across_shapes(obs, act, hidden, E):
    """The tensor-core kernels (rollout, pool rows, all-member forward, MOPO step, fused training
    forward) against the SIMT kernels on the same weights and inputs for other observation /
    action sizes, depths and widths than the MBPO default."""
    cfg = dict(E=E, elites=max(1, E - 1), hidden=hidden, act="swish")
    ms, _ = make_pe_model(cfg, 61, "cheetah", kernel="simt", obs=obs, act=act)
    mt, _ = make_pe_model(cfg, 61, "cheetah", kernel="tcgen05", obs=obs, act=act)
    if not mt._tc_supported():
        pytest.skip("shape not covered by the tile kernel")
    loss = list(np.linspace(0.9, 0.1, E))
    ms.remember_loss(loss)
    mt.remember_loss(loss)
    B = 777
    o, a = cases.rollout_inputs(62, B, obs=obs, act=act)
    o, a = T(o, "cuda"), T(a, "cuda")
    rs = np.random.RandomState(63)
    idx = rs.choice(mt.elite_indices, B)
    eps = T(rs.standard_normal((B, obs + 1)).astype(np.float32), "cuda")
    r1 = ms.step(o, a, indices=idx, noise=eps)
    r2 = mt.step(o, a, indices=idx, noise=eps)
    for x, y, what in zip(r1[:3], r2[:3], ("next_obs", "reward", "done")):
        assert_close(y, x, what=what)
    rows = torch.zeros(B, mt.row_stride(), device="cuda")
    mt.step_rows(o, a, rows, indices=idx, noise=eps, ordered=False)
    assert np.array_equal(_sorted_rows(rows[:, :mt.row_width()]), _sorted_rows(torch.cat([o, a, r2[0], r2[1], r2[2]], 1)))
    for x, y, what in zip(ms.predict_distribution(o, a), mt.predict_distribution(o, a), ("mean", "logstd")):
        assert_close(y, x, what=what)
    if E > 2:
        d1 = ms.step(o, a, indices=idx, noise=eps, return_distribution=True, penalty_type="total_var")
        d2 = mt.step(o, a, indices=idx, noise=eps, return_distribution=True, penalty_type="total_var")
        for k in d1[3]:
            assert_close(d2[3][k], d1[3][k], what=k)
    # fused training forward (scratch pack) against the per-layer path (no pack): same gradients
    bt = {k: T(v, "cuda") for k, v in cases.train_batch(64, E, 130, obs=obs, act=act).items()}
    from mirl_b200 import _lib
    Lb = _lib.lib()
    outs = []
    for pack in (None, torch.empty(Lb.mb_pe_tc_pack_bytes(ctypes.byref(mt._c_model(None))), dtype=torch.uint8, device="cuda")):
        cm = mt._c_model(pack)
        tcfg = mt.train_cfg(False)
        ws = torch.empty(Lb.mb_pe_train_workspace_bytes(ctypes.byref(cm), 130), dtype=torch.uint8, device="cuda")
        grads = torch.zeros_like(mt.module.flat.data)
        el = torch.zeros(E, device="cuda")
        terms = torch.zeros(3, device="cuda")
        _lib.check(Lb.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(tcfg), _lib.ptr(bt["observations"]),
                                       _lib.ptr(bt["actions"]), _lib.ptr(bt["deltas"]), _lib.ptr(bt["rewards"]),
                                       _lib.ptr(bt["terminals"]), 130, _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms),
                                       _lib.ptr(ws), ws.numel(), _lib.stream()))
        outs.append((grads.clone(), el.clone()))
    assert norm_rel(outs[1][0], outs[0][0]) < 1e-4
    assert_close(outs[1][1], outs[0][1], what="ensemble loss")
