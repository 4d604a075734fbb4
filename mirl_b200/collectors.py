"""Rollout drivers with the reference collectors' interface.

``B200MBPOCollector`` mirrors ``MBPOCollector`` (mirl/agents/mbpo/mbpo_collector.py:6-97) and
``B200MOPOCollector`` mirrors ``MOPOCollector`` (mirl/agents/mbpo/mopo_collector.py:7-68): same
constructor arguments, same horizon loop / stop mask / live-row compaction, same pool calls.
Differences: each ``model.step`` is one fused CUDA launch; MOPO's penalty is computed inside
that launch (``penalty_type`` is forwarded to ``B200PEModel.step``) instead of from six
materialised ``[n_elite,B,17]`` tensors; with ``torch.distributed`` initialised and
``shard=True`` the start states are partitioned by rows over the ranks and the generated
transitions are all-gathered before they enter the (replicated) pool.
"""
import numpy as np
import torch

from . import dist as mbdist
from .compat import Component


def _scheduled(current, schedule):
    """mirl/utils/misc_untils.py:5-12 get_scheduled_value."""
    if type(schedule) in (float, int):
        return schedule
    start, end, start_value, end_value = schedule
    ratio = max(0, min(1, (current - start) / (end - start)))
    return ratio * (end_value - start_value) + start_value


_KEYS = ("observations", "actions", "next_observations", "rewards", "terminals")


class B200MBPOCollector(Component):
    def __init__(self, model, policy, pool, imagined_data_pool, depth_schedule=1,
                 number_sample=int(1e5), save_n=20, bathch_size=100000, hard_stop=True,
                 rollout_terminal_state=False, step_kwargs={}, shard=False):
        self.model, self.policy, self.pool = model, policy, pool
        self.imagined_data_pool = imagined_data_pool
        self.depth_schedule = depth_schedule
        self.number_sample = int(number_sample)
        self.save_n = save_n
        self.batch_size = bathch_size if bathch_size is None else int(bathch_size)
        self.step_kwargs = dict(step_kwargs)
        self.hard_stop = hard_stop
        self.rollout_terminal_state = rollout_terminal_state
        self.shard = shard
        self._n_image_times_total = 0

    def imagine(self, cur_epoch):
        """mbpo_collector.py:34-54."""
        states = self.get_start_states()
        depth = self.cur_depth = self.get_depth(cur_epoch)
        self.resize_pool(cur_epoch)
        number_sample = states.shape[0]
        batch_size = self.batch_size or number_sample
        rank, ws = mbdist.world()
        sharded = self.shard and ws > 1
        with torch.no_grad():
            # loop bounds come from the GLOBAL row count: every rank runs the same number of chunks
            # (the loop body issues collectives when sharded) and takes its row slice of each chunk
            for i in range(0, number_sample, batch_size):
                o = states[i:i + min(batch_size, number_sample - i)]
                if sharded:
                    r0, r1 = mbdist.row_partition(o.shape[0], ws)[rank]
                    o = o[r0:r1]
                stop = torch.zeros((o.shape[0], 1), device=o.device)
                for j in range(depth):
                    a, _ = self.policy.action(o)
                    fused = self._fused_step(o, a)
                    if fused is not None:                             # rows already in the device pool
                        next_o, d = fused
                        if self.rollout_terminal_state:
                            o, stop = next_o, d
                        else:
                            # one host read decides everything: with no terminated row (the common case) the next
                            # batch IS next_o -- no mask, no gather; the survivors of a compaction all have stop == 0,
                            # so the loop ends exactly when no row is left (mbpo_collector.py:50-53,85-97)
                            # (d is a strided column of the pool rows: a compare + any() reads it an order of magnitude
                            # faster than a strided sum())
                            if bool(torch.any(d != 0).item()):
                                ind = self._get_live_ind(d)
                                next_o, d = next_o[ind], d[ind]
                            o, stop = next_o, d
                            if o.shape[0] == 0:
                                break
                            continue
                    else:
                        next_o, r, d, info = self.model.step(o, a, **self.step_kwargs)
                        d = stop + d - stop * d                       # stop or d
                        o, stop = self.add_sample(stop, o, a, next_o, r, d, info, j, cur_epoch)
                    if self._all_stopped(stop):
                        break
        self._n_image_times_total += 1

    def _fused_step(self, o, a):
        """Device-resident fast path: when the imagined pool is a ``B200DevicePool`` the rollout kernel
        writes this step's transitions straight into the pool's ring (``reserve`` + ``step_rows``),
        replacing ``model.step`` + five ``.cpu().numpy()`` copies + ``add_samples``
        (mbpo_collector.py:69-83).  Valid when every row of the batch is live, which the loop
        guarantees in the default configuration (terminated rows are dropped before the next step);
        returns ``None`` for the configurations that need the reference-shaped path."""
        from .pools import B200DevicePool
        pool = self.imagined_data_pool
        if (not isinstance(pool, B200DevicePool) or self.step_kwargs or self.rollout_terminal_state
                or (self.shard and mbdist.world()[1] > 1) or not hasattr(self.model, "step_rows")
                or pool.row_stride != self.model.row_stride() or o.shape[0] == 0):
            return None
        n = o.shape[0]
        start = pool.reserve(n)
        if start is None:
            # the block would wrap around the ring (or the pool keeps running statistics): the kernel fills a
            # staging block and the ring insert is one device copy (simple_pool.py:125-150 semantics)
            block = torch.empty(n, pool.row_stride, device=pool.rows.device)
            next_o, _, d = self.model.step_rows(o.contiguous(), a.contiguous(), block, row_offset=0, ordered=True)
            pool.add_rows(block)
            return next_o, d
        next_o, _, d = self.model.step_rows(o.contiguous(), a.contiguous(), pool.rows, row_offset=start, ordered=True)
        return next_o, d

    def _all_stopped(self, stop):
        n = torch.tensor([float(len(stop)), float(torch.sum(stop)) if len(stop) else 0.0],
                         device=stop.device)
        if self.shard and mbdist.world()[1] > 1:
            torch.distributed.all_reduce(n)      # every rank must leave the loop together
        return n[0].item() == 0 or n[1].item() == n[0].item()

    def get_start_states(self):
        """mbpo_collector.py:56-58.  With ``shard`` rank 0 draws the global batch and broadcasts it:
        the ranks' NumPy streams drift apart after the first call (each rank draws member indices
        for its own row count), so "every rank draws the same batch" cannot rely on them."""
        rank, ws = mbdist.world()
        dev = self.model.device
        n = int(self.number_sample)
        if self.shard and ws > 1:
            shape = [None]
            if rank == 0:
                states = self.pool.random_batch_torch(n, keys=["observations"])["observations"].to(dev).float()
                states = states.contiguous()
                shape = [tuple(states.shape)]
            torch.distributed.broadcast_object_list(shape, 0)
            if rank != 0:
                states = torch.empty(shape[0], device=dev)
            torch.distributed.broadcast(states, 0)
            return states
        data = self.pool.random_batch_torch(n, keys=["observations"])
        return data["observations"].to(dev)

    def _get_live_ind(self, stop):
        return (stop < 0.5).flatten()

    def _emit(self, samples):
        """D2H + pool insert (mbpo_collector.py:81-83); all-gather first when sharded."""
        if self.shard and mbdist.world()[1] > 1:
            widths = {k: int(np.prod(samples[k].shape[1:])) for k in _KEYS}       # also right for 0 rows
            packed = torch.cat([samples[k].reshape(samples[k].shape[0], widths[k]) for k in _KEYS], dim=1)
            packed = mbdist.all_gather_rows(packed.contiguous())
            out, c = {}, 0
            for k in _KEYS:
                w = widths[k]
                out[k] = packed[:, c:c + w]
                c += w
            samples = out
        if hasattr(self.imagined_data_pool, "add_rows"):               # device pool: no D2H / H2D round trip
            self.imagined_data_pool.add_samples(samples)
        else:
            self.imagined_data_pool.add_samples({k: v.cpu().numpy() for k, v in samples.items()})

    def _rewards(self, r, info):
        return r

    def add_sample(self, stop, o, a, next_o, r, d, info, depth, cur_epoch):
        """mbpo_collector.py:69-88."""
        ind = self._get_live_ind(stop)
        r = self._rewards(r, info)
        self._emit({"observations": o[ind], "actions": a[ind], "next_observations": next_o[ind],
                    "rewards": r[ind], "terminals": d[ind]})
        if self.rollout_terminal_state:
            return next_o, d
        ind = self._get_live_ind(d)
        return next_o[ind], d[ind]

    def get_depth(self, cur_epoch):
        self.depth = int(_scheduled(cur_epoch, self.depth_schedule))
        return self.depth

    def resize_pool(self, cur_epoch):
        size = int(self.depth * self.number_sample * self.save_n)
        if size != self.imagined_data_pool.max_size:
            self.imagined_data_pool.resize(size)


class B200MOPOCollector(B200MBPOCollector):
    def __init__(self, model, policy, pool, imagined_data_pool, penalty_coef=1,
                 penalty_type="total_var", **base_kwargs):
        super().__init__(model=model, policy=policy, pool=pool,
                         imagined_data_pool=imagined_data_pool, **base_kwargs)
        if penalty_type not in ("total_var", "max_var"):
            raise NotImplementedError(penalty_type)
        self.penalty_coef, self.penalty_type = penalty_coef, penalty_type
        self.step_kwargs["return_distribution"] = True                 # mopo_collector.py:26
        self.step_kwargs["penalty_type"] = penalty_type
        self.step_kwargs["penalty_coef"] = penalty_coef

    def _rewards(self, r, info):
        """mopo_collector.py:42-53: ``r - penalty_coef * penalty`` (computed in the step kernel)."""
        return info["penalized_reward"]
