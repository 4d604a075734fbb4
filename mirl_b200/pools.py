"""B200DevicePool -- the imagined-data pool of MBPO kept in HBM.

Drop-in for ``SimplePool`` (mirl/pools/simple_pool.py:15-184) as the ``imagined_data_pool`` of
``DynaStyleAlgorithm`` / ``MBPOCollector``: same fields, ring-buffer semantics (``add_samples``
wrap-around, simple_pool.py:125-150), ``random_batch`` index draw (``np.random.randint(0, size,
batch_size)`` from the global NumPy generator, pools/utils.py:37-39), ``get_data`` ordering,
``resize``, ``get_size``.  Differences: storage is ONE device tensor of packed rows
``[obs | act | next_obs | reward | done | pad]`` -- exactly what ``B200PEModel.step_rows`` writes, so
a rollout step inserts its transitions with no copy at all (``reserve`` + ``step_rows``), and
``random_batch_torch`` returns CUDA tensors without the per-update H2D copy of
``dyna_algorithm.py:115-124``.  ``compute_mean_std`` is not supported (the imagined pool never
uses it).
"""
import numpy as np
import torch

from .compat import Pool

_KEYS = ("observations", "actions", "next_observations", "rewards", "terminals")


class B200DevicePool(Pool):
    def __init__(self, env, max_size=1e6, compute_mean_std=False, device=None, row_stride=None, storage=None):
        if compute_mean_std:
            raise NotImplementedError("B200DevicePool: compute_mean_std is not supported")
        self.max_size = int(max_size)
        self.compute_mean_std = False
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
        if storage is not None:                 # e.g. PeerPools.local: a replica other GPUs write into
            assert storage.shape == (self.max_size, self.row_stride) and storage.is_cuda
            self.rows = storage
        else:
            self.rows = torch.zeros(self.max_size, self.row_stride, device=self.device)
        self._size = 0
        self._stop = 0

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
        if self._stop + n >= self.max_size:
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
        data = self.get_data_torch(self.sample_keys)
        block = torch.zeros(self._size, self.row_stride, device=self.device)
        for k, (c0, c1) in self.fields.items():
            block[:, c0:c1] = data[k]
        self.max_size = int(size)
        self.rows = torch.zeros(self.max_size, self.row_stride, device=self.device)
        self._size = 0
        self._stop = 0
        if block.shape[0]:
            self.add_rows(block)

    def get_diagnostics(self):
        from collections import OrderedDict
        d = OrderedDict([("size", self._size)])
        if self._size:
            t = self.get_data_torch(["terminals"])["terminals"]
            d["Done Ratio"] = float(t.sum()) / self._size
        return d
