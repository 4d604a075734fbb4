"""Two-process checks (launched by test_gpu_multi.py under torch.distributed.run, one rank per GPU):

1. ShardedPool: every rank keeps only the rows it generated; a seeded ``random_batch_rows`` must return
   exactly the rows a replicated pool (PeerPools filled by the same rollouts) returns for the same indices.
2. Member-sharded ``train_with_pool`` (members split over the ranks, global stop decision, weight
   all-gather) must end with the parameters, Adam-independent ranking and elite set of an unsharded run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirl_b200  # noqa: E402
from mirl_b200.dist import PeerPools, ShardedPool, member_partition  # noqa: E402
from mirl_b200.pools import B200ExtraFieldPool  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.helpers_gpu import FakeEnv, make_pe_model  # noqa: E402


def sharded_pool(rank, world, dev):
    m, _ = make_pe_model(cases.MBPO, 11, "hopper", device=dev)
    m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
    B, H, O = 2501 + 100 * rank, 3, m.obs_size            # ragged: the ranks hold different row counts
    S = m.row_stride()
    counts = [2501 + 100 * r for r in range(world)]
    rep = PeerPools(H * sum(counts), S, device=dev)
    shard = ShardedPool(H * max(counts), S, device=dev, seed=5)
    rep.local.fill_(-1.0)
    torch.cuda.synchronize()
    dist.barrier()
    rs = np.random.RandomState(77 + rank)
    o = torch.from_numpy(cases.rollout_inputs(500 + rank, B, env="hopper")[0]).to(dev)
    for h in range(H):
        a = torch.from_numpy(cases.rollout_inputs(600 + 10 * h + rank, B, env="hopper")[1]).to(dev)
        idx = rs.choice(m.elite_indices, B)
        eps = torch.from_numpy(rs.standard_normal((B, O + 1)).astype(np.float32)).to(dev)
        off = h * sum(counts) + sum(counts[:rank])               # replicated layout: [h][rank][row]
        m.step_rows(o, a, rep.destinations(), row_offset=off, indices=idx, noise=eps, ordered=True)
        start = shard.reserve(B, counts=counts if h else None)    # first block: counts all-gathered
        no, _, _ = m.step_rows(o, a, shard.local, row_offset=start, indices=idx, noise=eps, ordered=True)
        o = no.clone()
    rep.fence()
    shard.fence()
    torch.cuda.synchronize()
    assert len(shard) == H * sum(counts)
    gidx = shard.random_batch_index(4096)
    want_idx = np.random.RandomState(5).randint(0, len(shard), 4096)
    assert np.array_equal(gidx, want_idx)
    got = shard.rows(gidx)
    want = rep.local.index_select(0, torch.from_numpy(gidx).to(dev))
    ok = torch.equal(got, want)
    owner, _ = shard.locate(gidx)
    remote = int((owner != rank).sum())
    torch.cuda.synchronize()
    dist.barrier()
    shard.close()
    rep.close()
    return ok and remote > 1000


def sharded_training(rank, world, dev):
    c = cases.TRAINLOOP["small"]
    cfg = c["cfg"]
    res = []
    for members in (member_partition(cfg["E"], world)[rank], None):
        m, _ = make_pe_model(cfg, c["seed"], device=dev)
        env = FakeEnv()
        pool = B200ExtraFieldPool(env, max_size=5000, compute_mean_std=True, device=dev)
        pool.add_extra_fields({"deltas": {"shape": (cases.OBS,), "type": np.float32}})
        tr = mirl_b200.B200PEModelTrainer(env, m, lr=c["lr"], tb_log_freq=-1, silent=True, batch_size=c["batch_size"],
                                          report_freq=c["report_freq"], max_not_improve=c["max_not_improve"],
                                          max_valid=c["max_valid"], init_model_train_step=c["init_model_train_step"],
                                          max_step_per_train=c["max_step_per_train"], members=members)
        np.random.seed(c["rng_seed"])
        pool.add_samples(cases.dyn_samples(c["adds"][0][0], c["adds"][0][1]))
        st = tr.train_with_pool(pool)
        res.append((m.module.flat.data.clone(), list(m.elite_indices), st["train_step"]))
    (pa, ea, sa), (pb, eb, sb) = res
    return torch.equal(pa, pb) and ea == eb and sa == sb


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    ok1 = sharded_pool(rank, world, dev)
    print("rank %d sharded pool %s" % (rank, "OK" if ok1 else "MISMATCH"), flush=True)
    ok2 = sharded_training(rank, world, dev)
    print("rank %d sharded training %s" % (rank, "OK" if ok2 else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok1 and ok2 else 1)


if __name__ == "__main__":
    main()
