"""Multi-GPU partitioning of the path (one process per GPU, ``torch.distributed``).

* model training  -- partition by ENSEMBLE MEMBER: members are independent (the loss is a plain
  sum over members, pe_model.py:95-96; bootstrap batches are per member,
  pe_model_trainer.py:179,268), so the step needs no collective.  After training the ranks
  exchange their trained member slices (all-gather) and the E evaluation losses.
* rollout         -- partition by BATCH ROWS: every rank rolls its slice; the generated
  transitions are all-gathered into the (replicated) model pool (mbpo_collector.py:69-83 is the
  single-process equivalent).  Row counts differ per rank once terminations occur, so counts are
  gathered first and the payload is padded to the largest.
* critic targets  -- replicas only at batch 256 (latency bound); no collective.

Backends: ``nccl`` on GPUs (NVLink / NVSwitch), ``gloo`` in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def member_partition(ensemble_size, world_size):
    """Contiguous member slices, sizes differing by at most one: E=7 over 2 -> [(0,4),(4,7)];
    over 8 -> seven singletons and one empty slice."""
    base, rem = divmod(ensemble_size, world_size)
    out, e = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((e, e + n))
        e += n
    return out


def row_partition(n_rows, world_size):
    base, rem = divmod(n_rows, world_size)
    out, s = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def all_gather_members(module, partition=None, group=None):
    """After member-sharded training: every rank receives every member's parameters.
    ``module`` is a ``PackedEnsembleMLP``; ``partition[r]`` is rank r's member slice."""
    rank, ws = world()
    if ws == 1:
        return
    partition = partition or member_partition(module.ensemble_size, ws)
    per_member = sum(b - a for a, b in module.member_slices(0, 1))
    max_members = max(e1 - e0 for e0, e1 in partition)
    flat = module.flat.data
    send = torch.zeros(max_members * per_member, device=flat.device, dtype=flat.dtype)
    e0, e1 = partition[rank]
    if e1 > e0:
        send[:(e1 - e0) * per_member] = torch.cat([flat[a:b] for a, b in module.member_slices(e0, e1)])
    recv = torch.empty(ws * max_members * per_member, device=flat.device, dtype=flat.dtype)
    dist.all_gather_into_tensor(recv, send, group=group)
    for r, (a0, a1) in enumerate(partition):
        if a1 == a0 or r == rank:
            continue
        chunk = recv[r * max_members * per_member:][: (a1 - a0) * per_member]
        off = 0
        for a, b in module.member_slices(a0, a1):
            flat[a:b] = chunk[off:off + (b - a)]
            off += b - a
    module.mark_weights_changed()


def all_gather_vector(local, partition, group=None):
    """Gather a per-member vector (e.g. evaluation losses ``[E]``): rank r owns entries
    ``partition[r]``; everyone gets the full vector."""
    rank, ws = world()
    if ws == 1:
        return local
    mine = torch.zeros_like(local)
    e0, e1 = partition[rank]
    mine[e0:e1] = local[e0:e1]
    dist.all_reduce(mine, group=group)          # disjoint supports: sum == concatenation
    return mine


def all_gather_rows(t, group=None):
    """All-gather of a ``[n_r, C]`` tensor whose row count differs per rank -> ``[sum n_r, C]``
    in rank order.  Counts first, then one padded all-gather."""
    rank, ws = world()
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    counts = torch.empty(ws, device=t.device, dtype=torch.int64)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.tolist()
    m = max(counts)
    send = t.new_zeros((m,) + tuple(t.shape[1:]))
    send[: t.shape[0]] = t
    recv = t.new_empty((ws * m,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    return torch.cat([recv[r * m: r * m + c] for r, c in enumerate(counts)], dim=0)
