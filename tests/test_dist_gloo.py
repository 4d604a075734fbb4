"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: member-sharded weight exchange,
per-member vector gather, ragged all-gather of generated transitions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mirl_b200 import dist as mbd
        from mirl_b200.ensemble import PackedEnsembleMLP
        torch.manual_seed(0)                                    # same initial weights on every rank
        mod = PackedEnsembleMLP(23, 36, [32, 32], 7, activation="swish", probabilistic=True)
        reference = mod.flat.data.clone()
        part = mbd.member_partition(7, world)
        # every rank "trains" only its own members: member e becomes the constant e+1
        e0, e1 = part[rank]
        for e in range(e0, e1):
            for a, b in mod.member_slices(e, e + 1):
                mod.flat.data[a:b] = float(e + 1)
        mbd.all_gather_members(mod, part)
        ok = True
        for e in range(7):
            for a, b in mod.member_slices(e, e + 1):
                ok = ok and bool((mod.flat.data[a:b] == float(e + 1)).all())
        # per-member losses
        local = torch.zeros(7)
        local[e0:e1] = torch.arange(7.0)[e0:e1] * 10
        full = mbd.all_gather_vector(local, part)
        ok = ok and torch.equal(full, torch.arange(7.0) * 10)
        # ragged rows: rank r contributes r+2 rows
        t = torch.full((rank + 2, 5), float(rank))
        g = mbd.all_gather_rows(t)
        exp = torch.cat([torch.full((r + 2, 5), float(r)) for r in range(world)])
        ok = ok and torch.equal(g, exp)
        q.put((rank, ok, float(reference.sum())))
    finally:
        dist.destroy_process_group()


def test_member_and_row_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


class _Pool:
    max_size = 0

    def __init__(self):
        self.rows = []

    def resize(self, n):
        self.max_size = n

    def add_samples(self, s):
        self.rows.append({k: np.array(v) for k, v in s.items()})

    def random_batch_torch(self, n, keys=None):
        rs = np.random.RandomState(3)
        return {"observations": torch.from_numpy(rs.standard_normal((n, 4)).astype(np.float32))}


class _Model:
    device = torch.device("cpu")

    def step(self, o, a, **kw):
        nxt = o + 1.0
        d = (nxt[:, :1] > 2.0).float()                          # rows terminate at different depths
        return nxt, nxt[:, :1] * 0.5, d, {}


class _Policy:
    def action(self, o):
        return torch.zeros(o.shape[0], 2), {}


def test_collector_matches_reference_loop_semantics():
    """B200MBPOCollector.imagine reproduces mbpo_collector.py:34-88: every live row is stored,
    terminated rows are dropped from the next step, the loop exits when all rows stopped."""
    from mirl_b200.collectors import B200MBPOCollector
    pool, im = _Pool(), _Pool()
    c = B200MBPOCollector(_Model(), _Policy(), pool, im, depth_schedule=5, number_sample=64, bathch_size=32)
    c.imagine(0)
    assert im.max_size == 5 * 64 * 20
    start = pool.random_batch_torch(64)["observations"].numpy()
    # replay the reference loop in numpy
    exp_rows = 0
    for i in range(0, 64, 32):
        o = start[i:i + 32]
        for j in range(5):
            exp_rows += len(o)
            nxt = o + 1.0
            o = nxt[nxt[:, 0] <= 2.0]
            if len(o) == 0:
                break
    got = sum(len(r["observations"]) for r in im.rows)
    assert got == exp_rows
    assert set(im.rows[0].keys()) == {"observations", "actions", "next_observations", "rewards", "terminals"}


def _collector_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mirl_b200.collectors import B200MBPOCollector
        np.random.seed(100 + rank)                  # the ranks' NumPy streams are NOT in step
        pool, im = _Pool(), _Pool()
        # 65 rows in chunks of 32: chunks of 32, 32 and 1 rows -- the last chunk leaves rank 1 with no
        # rows at all, and the per-rank row counts straddle the chunk size
        c = B200MBPOCollector(_Model(), _Policy(), pool, im, depth_schedule=5, number_sample=65, bathch_size=32,
                              shard=True)
        c.imagine(0)
        got = sum(len(r["observations"]) for r in im.rows)
        first = np.concatenate([r["observations"] for r in im.rows])
        q.put((rank, got, float(first.sum())))
    finally:
        dist.destroy_process_group()


def test_sharded_collector_same_pool_on_every_rank_world2():
    """Row-sharded imagine(): loop bounds come from the global row count (every rank issues the same
    collectives even when a chunk leaves it without rows), rank 0's start-state draw is broadcast, and
    the all-gathered pool is the single-process pool on every rank."""
    from mirl_b200.collectors import B200MBPOCollector
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_collector_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pool, im = _Pool(), _Pool()
    c = B200MBPOCollector(_Model(), _Policy(), pool, im, depth_schedule=5, number_sample=65, bathch_size=32)
    c.imagine(0)
    want = sum(len(r["observations"]) for r in im.rows)
    want_sum = float(np.concatenate([r["observations"] for r in im.rows]).sum())
    for _, got, s in res:
        assert got == want
        assert abs(s - want_sum) < 1e-3 * abs(want_sum)
