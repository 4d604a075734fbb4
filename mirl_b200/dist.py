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


class _RawCudaArray:
    """``__cuda_array_interface__`` view of a raw device pointer (memory owned elsewhere)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerPools:
    """Replicated model pool filled by the rollout kernels themselves (north_star: "all-gather of
    generated transitions into the model pool"; SURVEY §8e prefers epilogue peer stores).

    Every rank owns one ``[rows, width]`` float32 replica (``.local``).  The replicas are exported
    over CUDA IPC and mapped into every other process, so ``destinations()`` -- the local tensor
    followed by the peers' mapped pointers -- can be handed to ``B200PEModel.step_rows``: each
    rank's kernel then writes its rows into ALL replicas while it computes, over NVLink, and the
    NCCL all-gather (which does not overlap with a persistent kernel) disappears.  A rank's rows
    are complete everywhere after its stream has drained and the ranks have passed ``fence()``.
    """

    def __init__(self, rows, width, device=None, group=None):
        import ctypes

        from . import _lib
        self._lib = _lib
        L = _lib.lib()
        self.rank, self.world = world()
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows, self.width = int(rows), int(width)
        nbytes = self.rows * self.width * 4
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(L.mb_peer_pool_alloc(nbytes, ctypes.byref(ptr), handle))
        self._own = ptr.value
        self.local = torch.as_tensor(_RawCudaArray(self._own, (self.rows, self.width)), device=self.device)
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = self._own
        self._mapped = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            with torch.cuda.device(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = ctypes.c_void_p()
                    _lib.check(L.mb_peer_pool_open(h, ctypes.byref(p)))
                    self.ptrs[r] = p.value
                    self._mapped.append(p.value)

    def destinations(self):
        """Local replica first (a tensor, so step_rows can return views into it), then the peers."""
        return [self.local] + [self.ptrs[r] for r in range(self.world) if r != self.rank]

    def fence(self):
        """All ranks' kernels issued so far have landed in every replica once this returns on the
        current stream's timeline (stream-ordered all-reduce of one element = barrier)."""
        if self.world > 1:
            if not hasattr(self, "_flag"):
                self._flag = torch.zeros(1, device=self.device)
            dist.all_reduce(self._flag, group=self.group)

    def close(self):
        L = self._lib.lib()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        for p in self._mapped:
            self._lib.check(L.mb_peer_pool_close(p))
        self._mapped = []
        if self.world > 1:
            dist.barrier(group=self.group)       # nobody frees while a peer still maps it
        if self._own is not None:
            self.local = None
            self._lib.check(L.mb_peer_pool_free(self._own))
            self._own = None
