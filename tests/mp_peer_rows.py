"""Two-process check of the fused compute + exchange path (launched by test_gpu_multi.py under
torch.distributed.run, one rank per GPU): every rank rolls its row slice with step_rows into ALL
pool replicas through peer-mapped stores; afterwards every replica must equal the concatenation of
what each rank's plain step() produced (exchanged here with an NCCL all-gather as the checker)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mirl_b200.dist import PeerPools  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.helpers_gpu import make_pe_model  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    for kernel in ("simt", "tcgen05"):
        m, _ = make_pe_model(cases.MBPO, 11, "hopper", device=dev, kernel=kernel)
        m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
        B, H, O, A = 3001, 3, m.obs_size, m.action_size
        W, S = m.row_width(), m.row_stride()
        ordered = kernel == "simt"          # tcgen05: bucket order = one bulk copy per tile and peer
        pools = PeerPools(H * world * B, S, device=dev)
        pools.local.fill_(-1.0)
        torch.cuda.synchronize()
        dist.barrier()
        o_np, _ = cases.rollout_inputs(500 + rank, B, env="hopper")
        rs = np.random.RandomState(77 + rank)
        o = torch.from_numpy(o_np).to(dev)
        o_ref = o.clone()
        want_local, o_ref_in = [], []
        for h in range(H):
            a = torch.from_numpy(cases.rollout_inputs(600 + 10 * h + rank, B, env="hopper")[1]).to(dev)
            idx = rs.choice(m.elite_indices, B)
            eps = torch.from_numpy(rs.standard_normal((B, O + 1)).astype(np.float32)).to(dev)
            o_ref_in.append(o_ref)
            no, r, d, _ = m.step(o_ref, a, indices=idx, noise=eps)
            want_local.append(torch.cat([o_ref, a, no, r, d, torch.zeros(B, S - W, device=dev)], 1))
            o_ref = no
            m.step_rows(o_ref_in[h], a, pools.destinations(), row_offset=(h * world + rank) * B,
                        indices=idx, noise=eps, ordered=ordered)
        pools.fence()
        torch.cuda.synchronize()
        mine = torch.stack(want_local)                                    # [H, B, W]
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = torch.stack(gathered, 1).reshape(H * world, B, S)          # [H][world] blocks of B rows
        got = pools.local.view(H * world, B, S)

        def srt(t):
            a_ = t.cpu().numpy()
            return a_[np.lexsort(a_.T[::-1])]
        ok = torch.equal(got, want) if ordered else all(
            np.array_equal(srt(got[j]), srt(want[j])) for j in range(H * world))
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        pools.close()
        if flag.item() != 1.0:
            print("rank %d kernel %s: replica mismatch" % (rank, kernel), flush=True)
            sys.exit(1)
        if rank == 0:
            print("peer rows OK kernel=%s world=%d" % (kernel, world), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
