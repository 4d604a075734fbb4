"""CPU tests of the host layer: C-ABI symbol surface, packed-parameter bookkeeping, reference
state-dict compatibility, RNG draw order, partition helpers.  No kernel is launched."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.golden import cases
from tests.helpers_gpu import FakeEnv
from tests.util import T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mirl_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "mirl_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", header))
    declared -= {"mb_mlp_desc", "mb_pe_model", "mb_train_cfg"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert L.mb_abi_version() == 1


def test_blob_layout_matches_c_offsets():
    from mirl_b200 import _lib
    from mirl_b200.ensemble import PackedEnsembleMLP
    m = PackedEnsembleMLP(23, 36, [200] * 4, 7, activation="swish", probabilistic=True)
    L = _lib.lib()
    off = 0
    for l in range(5):
        assert L.mb_mlp_weight_offset(m.desc, l) == off
        off = (off + 7 * m.layer_size[l] * m.layer_size[l + 1] + 3) & ~3
        assert L.mb_mlp_bias_offset(m.desc, l) == off
        off = (off + 7 * m.layer_size[l + 1] + 3) & ~3
    assert L.mb_mlp_logstd_offset(m.desc, 0) == off
    assert m.flat.numel() == L.mb_mlp_param_count(m.desc, 1)
    # views alias the blob
    m.weight(2)[3, 5, 7] = 42.0
    assert m.flat.data[m._w_off[2] + 3 * 200 * 200 + 5 * 200 + 7] == 42.0
    # member slices tile the trainable part of the blob exactly once
    covered = torch.zeros(m.flat.numel(), dtype=torch.int32)
    for e in range(7):
        for a, b in m.member_slices(e, e + 1):
            covered[a:b] += 1
    assert covered.max() == 1 and covered.sum() == 7 * (23 * 200 + 3 * 200 * 200 + 200 * 36 + 836 + 36)


def test_state_dict_speaks_the_reference_key_layout():
    import json
    import mirl_b200
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    m = mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish", connection="simple")
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref["PEModel"].keys())
    for k, shape in ref["PEModel"].items():
        assert list(sd[k].shape) == shape, k
    q = mirl_b200.B200EnsembleQValue(FakeEnv(), ensemble_size=10, hidden_layers=[256, 256], activation="relu")
    assert {k: list(v.shape) for k, v in q.state_dict().items()} == ref["EnsembleQValue"]
    t = mirl_b200.B200TQCQValue(FakeEnv(), ensemble_size=5, number_head=25, hidden_layers=[512] * 3, activation="relu")
    assert {k: list(v.shape) for k, v in t.state_dict().items()} == ref["TQCQValue"]


def test_state_dict_round_trip_and_per_member_files(tmp_path):
    import mirl_b200
    p = cases.pe_params(11, cases.MBPO)
    m = mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish", connection="simple")
    m.load_state_dict({k: T(v) for k, v in cases.pe_state_dict(p).items()})
    for l in range(5):
        assert np.array_equal(m.module.weight(l).numpy(), p["W"][l])
        assert np.array_equal(m.module.bias(l).numpy(), p["b"][l])
    assert np.array_equal(m.module.logstd_bound(0).numpy(), p["max_logstd"])
    assert np.array_equal(m.obs_processor.std.numpy(), p["obs_std"])
    # per-member snapshot files keyed by 'net{i}' (mlp.py:120-168)
    m.save(str(tmp_path), net_id=3)
    saved = torch.load(tmp_path / "PE_model_net3.pt")
    assert all("net3" in k for k in saved) and len(saved) == 12
    before = m.module.flat.data.clone()
    m.module.weight(1)[3].zero_()
    m.module.weight(1)[4].zero_()
    assert m.load(str(tmp_path), net_id=3)
    assert torch.equal(m.module.weight(1)[3], before[m.module._w_off[1]:].view(-1)[:7 * 200 * 200].view(7, 200, 200)[3])
    assert float(m.module.weight(1)[4].abs().sum()) == 0.0           # other members untouched
    with pytest.raises(RuntimeError):
        m.load_state_dict({"module.net.0.weight_net0": torch.zeros(3, 3)})


def test_elite_ranking_and_index_draw_follow_the_reference():
    import mirl_b200
    m = mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish", connection="simple")
    assert m.elite_indices is None
    m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
    assert list(m.elite_indices) == [1, 5, 3, 6, 0]                # pe_model.py:234-237
    g = np.load(os.path.join(ROOT, "tests", "golden", "pe_step_hopper.npz"))
    np.random.seed(22)
    idx = m._member_indices(1000, None).numpy()                       # draws np.random.choice like :169
    assert np.array_equal(idx.astype(np.int64), g["idx"])


def test_values_draw_member_subsets_like_the_reference():
    import mirl_b200
    from mirl_b200.values import sample_from_ensemble_indices
    g = np.load(os.path.join(ROOT, "tests", "golden", "redq.npz"))
    q = mirl_b200.B200EnsembleQValue(FakeEnv(), ensemble_size=10, hidden_layers=[256, 256], activation="relu")
    np.random.seed(73)
    members, rows, n = q._draw_subset(256, (256,), 2, True)
    assert np.array_equal(members, g["bw_idx"]) and rows is None and n == 2
    np.random.seed(74)
    raw = [np.random.choice(10 - i, (256,)) for i in range(2)]
    assert np.array_equal(np.stack(raw), g["row_raw_draws"])
    np.random.seed(74)
    table = sample_from_ensemble_indices(10, 256, 2, replace=False)
    assert (table[0] != table[1]).all() and table.max() < 10
    assert q._draw_subset(256, (256,), None, False)[:2] == (None, None)
    # parameters() yields per-member views in the reference order so soft updates can zip them
    ps = list(q.parameters())
    assert len(ps) == 60 and tuple(ps[0].shape) == (23, 256) and tuple(ps[1].shape) == (1, 256)
    ps[0].data.fill_(7.0)
    assert float(q.module.weight(0)[0].mean()) == 7.0


def test_unsupported_configurations_fail_loudly():
    import mirl_b200
    with pytest.raises(NotImplementedError):
        mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish", connection="densenet")
    with pytest.raises(NotImplementedError):
        mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish", known=["reward", "done"])
    with pytest.raises(NotImplementedError):
        mirl_b200.B200PEModel(FakeEnv("swimmer"), hidden_layers=[200] * 4, activation="swish")
    m = mirl_b200.B200PEModel(FakeEnv(), hidden_layers=[200] * 4, activation="swish")
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        m.step(torch.zeros(4, 17), torch.zeros(4, 6))               # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        mirl_b200.B200PEModelTrainer(FakeEnv(), m, optimizer_class="SGD")


def test_partitions():
    from mirl_b200.dist import member_partition, row_partition
    assert member_partition(7, 1) == [(0, 7)]
    assert member_partition(7, 2) == [(0, 4), (4, 7)]
    assert member_partition(7, 4) == [(0, 2), (2, 4), (4, 6), (6, 7)]
    assert member_partition(7, 8) == [(i, i + 1) for i in range(7)] + [(7, 7)]
    for n, w in ((100003, 8), (5, 8), (0, 2)):
        parts = row_partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_reference_factory_resolves_our_classes():
    """Container-only: with the reference importable, `"class": "B200PEModel"` in a MIRL config
    resolves through mirl/utils/initialize_utils.py:30-42 with no edit to the reference."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    import subprocess
    import sys
    code = (
        "from oracle import ref_shim; ref_shim.install()\n"
        "import mirl_b200\n"
        "from mirl_b200 import compat; assert compat.HAVE_REFERENCE\n"
        "from mirl.utils.initialize_utils import get_item_class\n"
        "assert get_item_class('component', 'B200PEModel') is mirl_b200.B200PEModel\n"
        "assert get_item_class('component', 'B200PEModelTrainer') is mirl_b200.B200PEModelTrainer\n"
        "assert get_item_class('component', 'B200MBPOCollector') is mirl_b200.B200MBPOCollector\n"
        "assert get_item_class('value', 'B200EnsembleQValue') is mirl_b200.B200EnsembleQValue\n"
        "assert get_item_class('value', 'B200TQCQValue') is mirl_b200.B200TQCQValue\n"
        "assert get_item_class('pool', 'B200DevicePool') is mirl_b200.B200DevicePool\n"
        "assert get_item_class('pool', 'B200ExtraFieldPool') is mirl_b200.B200ExtraFieldPool\n"
        "assert get_item_class('policy', 'B200GaussianPolicy') is mirl_b200.B200GaussianPolicy\n"
        "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on host cores) needs no GPU: one JSON line with
    the driver's keys, the reference-arm extras, and the same metric/config as our arm."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mbpo_rollout_transitions_per_s"
    assert line["unit"] == "transitions/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["value"] > 0 and line["steps"] == 1 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("MBPO model rollout")
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_sharded_pool_index_space_is_the_replicated_pools():
    """ShardedPool.locate: global row index -> (owner rank, local row), for ragged blocks, equals where an
    all-gather in rank order would have put the rows of every block (no GPU needed: the bookkeeping only)."""
    from mirl_b200.dist import ShardedPool
    world = 3
    pools = []
    for r in range(world):
        p = ShardedPool.__new__(ShardedPool)
        p._np, p.rank, p.world, p.capacity = np, r, world, 10_000
        p.reset()
        pools.append(p)
    rs = np.random.RandomState(0)
    blocks = [list(rs.randint(0, 50, world)) for _ in range(7)] + [[0, 5, 0]]
    owner_ref, local_ref = [], []
    used = [0] * world
    for counts in blocks:
        for r in range(world):
            assert pools[r].reserve(int(counts[r]), counts=[int(c) for c in counts]) == used[r]
        for r in range(world):                                   # all-gather order inside a block: rank 0, 1, 2
            owner_ref += [r] * counts[r]
            local_ref += list(range(used[r], used[r] + counts[r]))
            used[r] += counts[r]
    total = sum(sum(c) for c in blocks)
    assert len(pools[0]) == total == len(owner_ref)
    g = rs.permutation(total)
    for p in pools:
        owner, local = p.locate(g)
        assert np.array_equal(owner, np.asarray(owner_ref)[g]) and np.array_equal(local, np.asarray(local_ref)[g])
    with pytest.raises(RuntimeError):
        pools[0].reserve(20_000, counts=[20_000, 0, 0])


def test_actor_trainer_reduction_options():
    """B200ActorTrainer maps current_v_pi_kwargs onto the kernel's reduction like EnsembleQValue.value does
    (ensemble_value.py:65-96): all members, a batch-wise drawn subset from NumPy's global generator, and the options
    the chain does not cover raise instead of silently computing something else.  No GPU needed."""
    import numpy as np
    import mirl_b200
    from mirl_b200 import _lib
    from tests.helpers_gpu import FakeEnv
    env = FakeEnv()
    pol = mirl_b200.B200GaussianPolicy(env, hidden_layers=[32, 32], activation="relu")
    qf = mirl_b200.B200EnsembleQValue(env, ensemble_size=10, hidden_layers=[32, 32], activation="relu")

    def mk(kw):
        return mirl_b200.B200ActorTrainer(pol, qf, current_v_pi_kwargs=kw)
    assert mk({"sample_number": None, "mode": "mean"})._reduction() == (_lib.REDUCE["mean"], None)
    assert mk({"sample_number": 10})._reduction() == (_lib.REDUCE["min"], None)          # value()'s default mode
    np.random.seed(3)
    want = np.random.choice(10, 3, replace=False)
    np.random.seed(3)
    mode, members = mk({"sample_number": 3, "batchwise_sample": True, "mode": "max"})._reduction()
    assert mode == _lib.REDUCE["max"] and list(members) == list(want)
    with pytest.raises(NotImplementedError):
        mk({"sample_number": 2})._reduction()                                             # per-row draw
    with pytest.raises(NotImplementedError):
        mk({"mode": "sample", "sample_number": None})._reduction()
    tq = mirl_b200.B200TQCQValue(env, ensemble_size=5, number_head=25, hidden_layers=[32], activation="relu")
    assert mirl_b200.B200ActorTrainer(pol, tq, current_v_pi_kwargs={"number_drop": 0})._reduction() == \
        (_lib.REDUCE["mean"], None)
    with pytest.raises(NotImplementedError):
        mirl_b200.B200ActorTrainer(pol, tq, current_v_pi_kwargs={"number_drop": 5})._reduction()
    assert mirl_b200.B200ActorTrainer(pol, qf).target_entropy == -float(env.action_space.shape[0])   # sac_agent.py:55
