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


class ShardedPool:
    """Rank-sharded model pool ("keep the pool sharded ... with local sampling").

    ``PeerPools`` replicates every generated transition into all N pools, which costs every GPU
    (N-1) x 176 B of NVLink ingress per transition and caps the aggregate rollout rate at one GPU's ingress
    bandwidth whatever N is (measured: 3.98 G transitions/s at N = 8).  Here each rank keeps ONLY the rows
    it generated (its shard is still an IPC-exported buffer every peer maps), so a rollout step moves
    nothing over NVLink, and the exchange happens at SAMPLING time instead: a batch is drawn as global row
    indices -- the same on every rank, from a generator the ranks seed identically -- and the rows another
    rank owns are read through the mapping by the gather kernel itself (a 256-row SAC batch is 45 KB).

    The global index space is the one a replicated pool filled by an all-gather would have: rows are
    appended in BLOCKS (one per rollout step); inside a block rank 0's rows come first, then rank 1's, ...
    (``all_gather_rows`` order).  So a seeded ``random_batch_rows`` returns exactly the rows the replicated pool
    returns (tests/test_gpu_multi.py).  Append-only; ``reset()`` empties it (MBPO refills the imagined pool
    every ``imagine_freq`` loops)."""

    def __init__(self, rows_per_rank, width, device=None, group=None, seed=0):
        import numpy as np
        self._np = np
        self.peers = PeerPools(rows_per_rank, width, device=device, group=group)
        self.rank, self.world = self.peers.rank, self.peers.world
        self.device, self.width, self.capacity = self.peers.device, int(width), int(rows_per_rank)
        self.group = group
        self.local = self.peers.local
        # peer shards as tensors on THIS device (loads go over NVLink)
        self.views = [self.local if r == self.rank else
                      torch.as_tensor(_RawCudaArray(self.peers.ptrs[r], (self.capacity, self.width)), device=self.device)
                      for r in range(self.world)]
        self.rng = np.random.RandomState(seed)        # identical on every rank: the ranks' global streams drift
        self.reset()

    def reset(self):
        self.used = 0                                 # local rows in use
        self._starts = [0]                            # global start of every block (+ total at the end)
        self._counts = []                             # per block: rows per rank
        self._local0 = []                             # per block: first local row on every rank

    def __len__(self):
        return self._starts[-1]

    def reserve(self, n_local, counts=None):
        """Claim ``n_local`` rows of this rank's shard for a block whose rows the caller writes in place
        (``step_rows(..., dest=pool.local, row_offset=...)``).  ``counts`` = rows of every rank in this block
        when they are known (fixed-size rollouts); otherwise they are all-gathered (8 bytes per rank).
        Every rank must call this for every block.  Returns the local row offset."""
        np = self._np
        if counts is None:
            if self.world > 1:
                c = torch.tensor([n_local], device=self.device, dtype=torch.int64)
                out = torch.empty(self.world, device=self.device, dtype=torch.int64)
                dist.all_gather_into_tensor(out, c, group=self.group)
                counts = out.tolist()
            else:
                counts = [n_local]
        assert counts[self.rank] == n_local
        if self.used + n_local > self.capacity:
            raise RuntimeError("ShardedPool: shard full (%d + %d > %d rows)" % (self.used, n_local, self.capacity))
        prev = self._local0[-1] + np.asarray(self._counts[-1]) if self._counts else np.zeros(self.world, dtype=np.int64)
        self._local0.append(np.asarray(prev, dtype=np.int64))
        self._counts.append(np.asarray(counts, dtype=np.int64))
        self._starts.append(self._starts[-1] + int(sum(counts)))
        start = self.used
        self.used += n_local
        return start

    def locate(self, gidx):
        """Global row indices -> (owner rank, local row) arrays."""
        np = self._np
        gidx = np.asarray(gidx, dtype=np.int64)
        starts = np.asarray(self._starts, dtype=np.int64)
        blk = np.searchsorted(starts, gidx, side="right") - 1
        off = gidx - starts[blk]
        counts = np.stack(self._counts)[blk]                         # [n, world]
        cum = np.cumsum(counts, axis=1)
        owner = (off[:, None] >= cum).sum(1)
        before = np.where(owner > 0, cum[np.arange(len(gidx)), np.maximum(owner - 1, 0)], 0)
        local = np.stack(self._local0)[blk, owner] + (off - before)
        return owner, local

    def fence(self):
        """Every rank's rollout kernels issued so far have finished writing their shards (needed before
        a rank samples rows another rank generated): one stream-ordered 4-byte all-reduce."""
        self.peers.fence()

    def random_batch_index(self, batch_size):
        return self.rng.randint(0, len(self), batch_size)           # pools/utils.py:38 on the shared generator

    def rows(self, gidx):
        """Gather rows by global index: local rows from HBM, the others through the peer mappings."""
        owner, local = self.locate(gidx)
        out = torch.empty(len(owner), self.width, device=self.device)
        for r in range(self.world):
            sel = self._np.nonzero(owner == r)[0]
            if len(sel) == 0:
                continue
            li = torch.from_numpy(local[sel]).to(self.device)
            out[torch.from_numpy(sel).to(self.device)] = self.views[r].index_select(0, li)
        return out

    def random_batch_rows(self, batch_size):
        return self.rows(self.random_batch_index(batch_size))

    def close(self):
        self.views = None
        self.local = None
        self.peers.close()
