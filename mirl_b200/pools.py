"""B200DevicePool -- the imagined-data pool of MBPO kept in HBM.

Drop-in for ``SimplePool`` (mirl/pools/simple_pool.py:15-184) as the ``imagined_data_pool`` of
``DynaStyleAlgorithm`` / ``MBPOCollector``: same fields, ring-buffer semantics (``add_samples``
wrap-around, simple_pool.py:125-150), ``random_batch`` index draw (``np.random.randint(0, size,
batch_size)`` from the global NumPy generator, pools/utils.py:37-39), ``get_data`` ordering,
``resize``, ``get_size``.  Differences: storage is ONE device tensor of packed rows
``[obs | act | next_obs | reward | done | pad]`` -- exactly what ``B200PEModel.step_rows`` writes, so
a rollout step inserts its transitions with no copy at all (``reserve`` + ``step_rows``), and
``random_batch_torch`` returns CUDA tensors without the per-update H2D copy of
``dyna_algorithm.py:115-124``.

``B200ExtraFieldPool`` is the real-data pool of the same algorithms (``ExtraFieldPool`` over
``FlagPool``, mirl/pools/extra_field_pool.py:6-71, flag_pool.py:8-66) kept in HBM as well: the
running mean/std of every field (``RunningMeanStd``, mirl/utils/mean_std.py:3-33, float64, merged
batch by batch with the parallel-variance formula), extra fields such as the ``deltas`` the model
trainer maintains (pe_model_trainer.py:140-144), and the ``compute_delta`` processed-flag protocol.
With it ``B200PEModelTrainer.train_with_pool`` never touches host NumPy arrays.
"""
import numpy as np
import torch

from .compat import Pool

_KEYS = ("observations", "actions", "next_observations", "rewards", "terminals")


class DeviceRunningMeanStd(object):
    """``RunningMeanStd`` (mirl/utils/mean_std.py:3-33) with the float64 state on the device.
    The per-batch moments are reduced in float64 (the reference's ``np.mean`` / ``np.var`` of a
    float32 batch reduce in float32 pairwise sums: same value to ~1e-7 relative)."""

    def __init__(self, shape=(), device=None, epsilon=1e-4):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.std = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = epsilon
        self.epsilon = epsilon

    def update(self, x):
        x = x.double()
        n = x.shape[0]
        batch_mean = x.mean(0)
        batch_var = x.var(0, unbiased=False)
        delta = batch_mean - self.mean                       # mean_std.py:22-33
        tot = self.count + n
        new_mean = self.mean + delta * n / tot
        m2 = self.var * self.count + batch_var * n + delta * delta * self.count * n / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot
        self.std = torch.sqrt(self.var)


class B200DevicePool(Pool):
    def __init__(self, env, max_size=1e6, compute_mean_std=False, device=None, row_stride=None, storage=None):
        self.max_size = int(max_size)
        self.compute_mean_std = bool(compute_mean_std)
        self._observation_shape = env.observation_space.shape
        self._action_shape = env.action_space.shape
        O, A = self._observation_shape[0], self._action_shape[0]
        self._O, self._A = O, A
        self.row_width = 2 * O + A + 2
        self.row_stride = int(row_stride) if row_stride else (self.row_width + 3) // 4 * 4
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # column ranges in SimplePool's field order (simple_pool.py:57-81)
        self.fields = {
            "observations": (0, O), "next_observations": (O + A, 2 * O + A), "actions": (O, O + A),
            "rewards": (2 * O + A, 2 * O + A + 1), "terminals": (2 * O + A + 1, 2 * O + A + 2),
        }
        self.sample_keys = list(self.fields.keys())
        self._external_storage = storage is not None
        if storage is not None:                 # e.g. PeerPools.local: a replica other GPUs write into
            assert storage.shape == (self.max_size, self.row_stride) and storage.is_cuda
            self.rows = storage
        else:
            self.rows = torch.zeros(self.max_size, self.row_stride, device=self.device)
        self._size = 0
        self._stop = 0
        if self.compute_mean_std:
            self.dataset_mean_std = {k: DeviceRunningMeanStd((c1 - c0,), self.device)
                                     for k, (c0, c1) in self.fields.items()}

    # ---- SimplePool interface ------------------------------------------------------------------
    def get_size(self):
        return self._size

    def _check_keys(self, keys, without_keys):
        if keys is None:
            keys = [k for k in self.fields.keys() if k not in without_keys]
        for k in keys:
            assert k in self.fields
        return keys

    def _pack(self, samples):
        n = len(samples["observations"])
        block = torch.zeros(n, self.row_stride, device=self.device)
        for k, (c0, c1) in self.fields.items():
            v = samples[k]
            v = torch.as_tensor(np.asarray(v), device=self.device) if not torch.is_tensor(v) else v.to(self.device)
            block[:, c0:c1] = v.reshape(n, c1 - c0).float()
        return block

    def add_rows(self, block):
        """Ring insert of packed rows ``[n, row_stride]`` (simple_pool.py:125-150 for every field at once)."""
        n = block.shape[0]
        if self.compute_mean_std:
            for k, (c0, c1) in self.fields.items():
                self.dataset_mean_std[k].update(block[:, c0:c1])
        stop, max_size = self._stop, self.max_size
        new_stop = (stop + n) % max_size
        if stop + n >= max_size:
            self.rows[stop:max_size] = block[:max_size - stop]
            if new_stop:
                self.rows[:new_stop] = block[n - new_stop:]
        else:
            self.rows[stop:new_stop] = block
        self._stop = new_stop
        self._size = min(max_size, self._size + n)
        return n

    def add_samples(self, samples):
        return self.add_rows(self._pack(samples))

    def reserve(self, n):
        """Claim the next ``n`` rows for a writer that fills them in place (``step_rows`` with
        ``row_offset``): returns the first row, or ``None`` when the block would wrap around the ring
        (use ``add_rows`` on a staging block then).  The pool counts the rows as present."""
        if self._stop + n >= self.max_size or self.compute_mean_std:
            return None
        start = self._stop
        self._stop += n
        self._size = min(self.max_size, self._size + n)
        return start

    def random_batch_index(self, batch_size):
        return np.random.randint(0, self._size, batch_size)             # pools/utils.py:38

    def random_batch_torch(self, batch_size, keys=None, without_keys=[]):
        keys = self._check_keys(keys, without_keys)
        idx = torch.from_numpy(self.random_batch_index(batch_size)).to(self.device)
        rows = self.rows.index_select(0, idx)
        return {k: rows[:, self.fields[k][0]:self.fields[k][1]] for k in keys}

    def random_batch(self, batch_size, keys=None, without_keys=[]):
        return {k: v.cpu().numpy() for k, v in self.random_batch_torch(batch_size, keys, without_keys).items()}

    def get_data_torch(self, keys=None, without_keys=[]):
        keys = self._check_keys(keys, without_keys)
        if self._size < self.max_size:
            rows = self.rows[:self._size]
        else:                                                            # simple_pool.py:118-120
            stop = self._stop                                            # full ring: oldest row sits at `stop`
            rows = torch.cat((self.rows[stop:], self.rows[:stop])) if stop else self.rows
        return {k: rows[:, self.fields[k][0]:self.fields[k][1]] for k in keys}

    def get_data(self, keys=None, without_keys=[]):
        return {k: v.cpu().numpy() for k, v in self.get_data_torch(keys, without_keys).items()}

    def resize(self, size, **init_kwargs):
        """simple_pool.py:175-184: re-create with the new capacity and re-insert the data in order."""
        assert size > self._size
        if self._external_storage:
            # the rows live in a buffer other GPUs write into (dist.PeerPools): replacing it here would
            # silently detach this pool from the peers' mapping
            raise RuntimeError("B200DevicePool.resize: the pool was built on external storage; "
                               "re-create the PeerPools with the new capacity instead")
        data = self.get_data_torch(self.sample_keys)
        block = torch.zeros(self._size, self.row_stride, device=self.device)
        for k, (c0, c1) in self.fields.items():
            block[:, c0:c1] = data[k]
        self.max_size = int(size)
        self.rows = torch.zeros(self.max_size, self.row_stride, device=self.device)
        self._size = 0
        self._stop = 0
        if self.compute_mean_std:               # simple_pool.py:183: __init__ again, then add_samples
            self.dataset_mean_std = {k: DeviceRunningMeanStd((c1 - c0,), self.device)
                                     for k, (c0, c1) in self.fields.items()}
        if block.shape[0]:
            self.add_rows(block)

    def get_mean_std(self, keys=None, without_keys=[]):
        """simple_pool.py:96-104: ``{key: [mean, std]}`` as float64 NumPy arrays (a few hundred
        bytes of D2H; the statistics themselves are accumulated on the device)."""
        assert self.compute_mean_std
        keys = self._check_keys(keys, without_keys)
        return {k: [self.dataset_mean_std[k].mean.cpu().numpy(), self.dataset_mean_std[k].std.cpu().numpy()]
                for k in keys}

    def get_diagnostics(self):
        from collections import OrderedDict
        d = OrderedDict([("size", self._size)])
        if self._size:
            t = self.get_data_torch(["terminals"])["terminals"]
            d["Done Ratio"] = float(t.sum()) / self._size
        return d


class B200ExtraFieldPool(B200DevicePool):
    """``ExtraFieldPool`` (extra_field_pool.py:6-71) over ``FlagPool`` (flag_pool.py:8-66), in HBM."""

    def __init__(self, env, max_size=1e6, extra_fields={}, compute_mean_std=False, device=None):
        super().__init__(env, max_size, compute_mean_std, device=device)
        self.unprocessed_size = {}
        self.extra_fields = {}
        self.extra_data = {}
        self.required_samples = {}
        self.extra_fields_stop = {}
        self.extra_fields_size = {}
        self.add_extra_fields(extra_fields)

    @property
    def extra_keys(self):
        return list(self.extra_fields.keys())

    def add_extra_fields(self, extra_fields):
        for k, v in extra_fields.items():
            assert k not in self.fields
            self.required_samples[k] = self._size
            self.extra_fields_stop[k] = 0
            self.extra_fields_size[k] = 0
            shape = tuple(v["shape"])
            self.extra_data[k] = torch.zeros((self.max_size,) + shape, device=self.device)
            if self.compute_mean_std:
                self.dataset_mean_std[k] = DeviceRunningMeanStd(shape, self.device)
        self.extra_fields.update(extra_fields)

    def add_rows(self, block):
        n = super().add_rows(block)
        for tag in self.unprocessed_size:                                   # flag_pool.py:24-30
            self.unprocessed_size[tag] = min(self.max_size, self.unprocessed_size[tag] + n)
        for key in self.extra_fields:                                       # extra_field_pool.py:33-36
            self.required_samples[key] += n
        return n

    def reserve(self, n):
        return None            # rows always enter through add_rows (flags and statistics)

    def _check_keys(self, keys, without_keys):
        if keys is None:
            keys = [k for k in list(self.fields.keys()) + list(self.extra_fields.keys())
                    if k not in without_keys]
        for k in keys:
            if k in self.extra_fields:
                if self.required_samples[k] > 0:
                    raise RuntimeError("the %s data is not aligned with the sampled data." % (k))
            else:
                assert k in self.fields
        return keys

    def _ring(self, t, size=None):
        size = self._size if size is None else size
        if size < self.max_size:
            return t[:size]
        stop = self._stop
        return torch.cat((t[stop:], t[:stop])) if stop else t

    def get_data_torch(self, keys=None, without_keys=[]):
        keys = self._check_keys(keys, without_keys)
        rows = None
        out = {}
        for k in keys:
            if k in self.fields:
                if rows is None:
                    rows = self._ring(self.rows)
                out[k] = rows[:, self.fields[k][0]:self.fields[k][1]]
            else:
                out[k] = self._ring(self.extra_data[k])
        return out

    def random_batch_torch(self, batch_size, keys=None, without_keys=[]):
        keys = self._check_keys(keys, without_keys)
        idx = torch.from_numpy(self.random_batch_index(batch_size)).to(self.device)
        rows = self.rows.index_select(0, idx)
        return {k: (rows[:, self.fields[k][0]:self.fields[k][1]] if k in self.fields
                    else self.extra_data[k].index_select(0, idx)) for k in keys}

    def get_unprocessed_data_torch(self, tag, keys=None, without_keys=[]):
        """flag_pool.py:33-48, device tensors."""
        keys = self._check_keys(keys, without_keys)
        if tag not in self.unprocessed_size:
            self.unprocessed_size[tag] = self._size
        stop, size = self._stop, self.unprocessed_size[tag]
        data = {}
        for key in keys:
            assert key in self.fields
            c0, c1 = self.fields[key]
            if size > stop:
                data[key] = torch.cat((self.rows[self.max_size + stop - size:, c0:c1], self.rows[:stop, c0:c1]))
            else:
                data[key] = self.rows[stop - size:stop, c0:c1]
        return data

    def get_unprocessed_data(self, tag, keys=None, without_keys=[]):
        return {k: v.cpu().numpy() for k, v in self.get_unprocessed_data_torch(tag, keys, without_keys).items()}

    def update_process_flag(self, tag, process_num):
        if tag not in self.unprocessed_size:
            self.unprocessed_size[tag] = self._size
        assert process_num <= self.unprocessed_size[tag]
        self.unprocessed_size[tag] -= process_num

    def update_single_extra_field(self, key, value):
        """extra_field_pool.py:47-66."""
        assert key in self.extra_fields
        value = value.to(self.device).float() if torch.is_tensor(value) else \
            torch.as_tensor(np.asarray(value), device=self.device).float()
        if self.compute_mean_std:
            self.dataset_mean_std[key].update(value)
        n, max_size = len(value), self.max_size
        assert n <= self.required_samples[key], "%d, %d" % (n, self.required_samples[key])
        self.required_samples[key] -= n
        stop = self.extra_fields_stop[key]
        self.extra_fields_stop[key] = new_stop = (stop + n) % max_size
        self.extra_fields_size[key] = min(max_size, self.extra_fields_size[key] + n)
        buf = self.extra_data[key]
        if stop + n >= max_size:
            buf[stop:max_size] = value[:max_size - stop]
            if new_stop:
                buf[:new_stop] = value[n - new_stop:]
        else:
            buf[stop:new_stop] = value

    def update_extra_fields(self, data):
        for k, v in data.items():
            self.update_single_extra_field(k, v)

    def resize(self, size, **init_kwargs):
        """extra_field_pool.py:73-79 over flag_pool.py:56-66."""
        extra = self.get_data_torch(self.extra_keys)
        extra = {k: v.clone() for k, v in extra.items()}
        for k in self.required_samples:
            assert self.required_samples[k] == 0
        flags = dict(self.unprocessed_size)
        for tag in flags:
            assert flags[tag] == 0
        fields = dict(self.extra_fields)
        self.unprocessed_size = {}
        self.extra_fields, self.extra_data = {}, {}
        self.required_samples, self.extra_fields_stop, self.extra_fields_size = {}, {}, {}
        super().resize(size)
        self.add_extra_fields(fields)
        for k in fields:                         # the rows were re-inserted before the fields existed
            self.required_samples[k] = self._size
        self.update_extra_fields(extra)
        self.unprocessed_size = flags
