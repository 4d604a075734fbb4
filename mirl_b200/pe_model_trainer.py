"""B200PEModelTrainer -- drop-in for ``PEModelTrainer``
(mirl/agents/mbpo/pe_model_trainer.py:18-306).

``train_from_torch_batch`` is ONE persistent kernel launch (forward + Gaussian NLL + log-std
bound + L2 + backward + Adam, ``mb_pe_train_step`` -> ``train_fused.cu``) instead of ~4 500 ATen
ops.  ``sufficiently_train`` keeps the reference's control flow (bootstrap
indices, periodic hold-out evaluation, per-member early stopping, best-snapshot restore, elite
ranking) and its NumPy RNG consumption order, but keeps the dataset, the bootstrap index table
and the per-member "best" snapshots in device memory instead of NumPy fancy-indexing + H2D every
step (:268-270) and ``torch.save`` per improved member (:249,256) -- and runs every block of
``report_freq`` steps between two evaluations as ONE launch (``mb_pe_train_steps``: the kernel
gathers each step's bootstrap batch from the index table itself).  With a ``B200ExtraFieldPool``
``train_with_pool`` never touches host NumPy arrays (the permutation / bootstrap draws still
come from the global NumPy generator, in the reference's order).

Member-sharded training (``members=(e0, e1)``): members are independent (pe_model.py:95-96), so
each rank trains its slice with no collective in the step; ``mirl_b200.dist.all_gather_members``
exchanges the trained slices afterwards.
"""
import ctypes
import os
import os.path as osp
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .compat import Component, logger, snapshot_dir


def split_dataset(dataset, valid_ratio, max_valid=5000):
    """mirl/pools/utils.py:60-79 (same ``np.random.permutation`` consumption)."""
    n_sample = len(next(iter(dataset.values())))
    valid_length = int(np.ceil(min(max_valid, valid_ratio * n_sample)))
    train_length = n_sample - valid_length
    indices = np.random.permutation(n_sample)
    train_indices, valid_indices = indices[:train_length], indices[-valid_length:]
    train = {k: v[train_indices] for k, v in dataset.items()}
    valid = {k: v[valid_indices] for k, v in dataset.items()}
    return train, valid, train_length, valid_length


_FIELDS = ("observations", "actions", "deltas", "rewards", "terminals")


class B200PEModelTrainer(Component):
    def __init__(self, env, model, lr=1e-3, max_step_per_train=4000, init_model_train_step=int(5e4),
                 load_best=True, valid_ratio=0.2, max_valid=2000, resample=True, batch_size=256,
                 report_freq=100, max_not_improve=20, silent=False, load_dataset_dir=None,
                 tb_log_freq=100, optimizer_class="Adam", optimizer_kwargs={}, load_dir=None,
                 save_dir=None, ignore_terminal_state=False, members=None, snapshot_to_disk=False,
                 use_cuda_graph=False):
        if optimizer_class not in ("Adam", torch.optim.Adam):
            raise NotImplementedError("the fused training step implements torch.optim.Adam only")
        extra = set(optimizer_kwargs) - {"betas", "eps"}
        if extra or optimizer_kwargs.get("weight_decay", 0) or optimizer_kwargs.get("amsgrad", False):
            raise NotImplementedError("Adam options %s" % sorted(extra))
        self.env = env
        self.model = model
        self.lr = lr
        self.optimizer_kwargs = dict(optimizer_kwargs)
        self.betas = tuple(optimizer_kwargs.get("betas", (0.9, 0.999)))
        self.eps = optimizer_kwargs.get("eps", 1e-8)
        flat = model.module.flat
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.statistics = OrderedDict()
        self.learn_reward, self.learn_done = model.learn_reward, model.learn_done
        self.num_train_steps = 0
        self.max_step_per_train = max_step_per_train
        self.init_model_train_step = int(init_model_train_step)
        self.load_best = load_best
        self.tb_log_freq = tb_log_freq
        self.ignore_terminal_state = ignore_terminal_state
        self.valid_ratio, self.max_valid, self.resample = valid_ratio, max_valid, resample
        self.batch_size, self.report_freq = batch_size, report_freq
        self.max_not_improve, self.silent = max_not_improve, silent
        self.load_dataset_dir = load_dataset_dir
        self.load_dir = osp.expanduser(load_dir) if load_dir is not None else None
        self.save_dir = osp.expanduser(save_dir) if save_dir is not None else None
        self.members = tuple(members) if members is not None else (0, model.ensemble_size)
        self.snapshot_to_disk = snapshot_to_disk
        self.use_cuda_graph = use_cuda_graph
        self._graph = None            # (key, CUDAGraph, static batch dict)
        self._step_dev = None         # device-resident Adam step counter (graph replay)
        self._ws = None
        self._loss_terms = None
        self._ensemble_loss = None
        self._best = None
        self._pack = None

    # ---- fused step ------------------------------------------------------------------------------
    def _state_to(self, dev):
        if self.exp_avg.device != dev:
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(dev), self.exp_avg_sq.to(dev)
        if self._ensemble_loss is None or self._ensemble_loss.device != dev:
            self._ensemble_loss = torch.zeros(self.model.ensemble_size, device=dev)
            self._loss_terms = torch.zeros(3, device=dev)

    def _scratch_pack(self):
        """Scratch operand pack for the fused tcgen05 forward of the training step (re-packed from the
        current weights inside every step); None when the tile kernel does not cover the shape."""
        model, L = self.model, _lib.lib()
        if not model._tc_supported():
            return None
        dev = model.device
        if self._pack is None or self._pack.device != dev:
            m = model._c_model(None)
            self._pack = torch.empty(L.mb_pe_tc_pack_bytes(ctypes.byref(m)), dtype=torch.uint8, device=dev)
        return self._pack

    def _launch(self, t, B, step_counter):
        """Enqueue the fused step on the current stream (one C-ABI call, ~10 kernel launches)."""
        model, L = self.model, _lib.lib()
        m = model._c_model(self._scratch_pack())
        cfg = model.train_cfg(self.ignore_terminal_state, self.lr, self.betas, self.eps)
        e0, e1 = self.members
        _lib.check(L.mb_pe_train_step(
            ctypes.byref(m), ctypes.byref(cfg), _lib.ptr(model.module.flat.data),
            _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.num_train_steps,
            _lib.ptr(step_counter) if step_counter is not None else None,
            _lib.ptr(t["observations"]), _lib.ptr(t["actions"]), _lib.ptr(t["deltas"]),
            _lib.ptr(t["rewards"]), _lib.ptr(t["terminals"]), B, e0, e1, _lib.ptr(self._ensemble_loss),
            _lib.ptr(self._loss_terms), _lib.ptr(self._ws), self._ws.numel(), _lib.stream()))

    @_lib.on_device
    def train_step_async(self, batch):
        """Enqueue one fused training step; returns device tensors
        ``(ensemble_loss [E], loss_terms [3] = sum_e loss_e, l2, bound)`` without synchronising
        (the reference syncs twice per step, pe_model_trainer.py:101-104).

        The step is ~10 kernels launched from ONE C call with programmatic dependent launch (each
        kernel's set-up runs under its predecessor's tail, also across consecutive steps).  With
        ``use_cuda_graph`` the sequence is captured once per batch shape and replayed instead (the
        Adam step count then lives on the device); measured slower on B200 (0.148 vs 0.132 ms):
        a replay copies the five input tensors into the graph's static buffers and consecutive
        replays are fully serialised."""
        model, L = self.model, _lib.lib()
        dev = model.device
        self._state_to(dev)
        t = {k: _lib.f32(batch[k], dev) for k in _FIELDS}
        o = t["observations"]
        assert o.dim() == 3 and o.shape[0] == model.ensemble_size, \
            "training batches are per-member [E,B,.] (pe_model_trainer.py:268-270)"
        B = o.shape[1]
        m = model._c_model(None)
        need = L.mb_pe_train_workspace_bytes(ctypes.byref(m), B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._graph = None
        self.num_train_steps += 1
        if not self.use_cuda_graph:
            self._launch(t, B, None)
        else:
            if self._step_dev is None or self._step_dev.device != dev:
                self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._scratch_pack()                     # allocated outside the capture
            key = (B, model.module.flat.data_ptr(), model._norm().data_ptr(), self.ignore_terminal_state,
                   self.lr, self.members, self.betas, self.eps, tuple(model.weight_decay),
                   model.reward_coef, model.bound_reg_coef, self.exp_avg.data_ptr(),
                   self.exp_avg_sq.data_ptr(), self._ws.data_ptr(), self._ensemble_loss.data_ptr(),
                   self._pack.data_ptr() if self._pack is not None else 0)
            if self._graph is None or self._graph[0] != key:
                static = {k: torch.empty_like(v) for k, v in t.items()}
                self._step_dev.fill_(self.num_train_steps - 1)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                # capture does not execute: the first replay below performs step `num_train_steps`
                with torch.cuda.graph(g, stream=side):
                    self._launch(static, B, self._step_dev)
                self._graph = (key, g, static)
            _, g, static = self._graph
            for k in _FIELDS:
                static[k].copy_(t[k], non_blocking=True)
            g.replay()
        model.module.mark_weights_changed()
        return self._ensemble_loss, self._loss_terms

    def fused_steps_supported(self):
        L = _lib.lib()
        m = self.model._c_model(None)
        return L.mb_pe_train_steps_workspace_bytes(ctypes.byref(m), 1) > 0

    def train_steps_async(self, data, index, train_length, batch_size, first_step, nsteps):
        """``nsteps`` consecutive training steps of the epoch schedule in ONE persistent launch
        (pe_model_trainer.py:240,268-272 for steps ``first_step .. first_step + nsteps - 1``): ``data``
        holds the device dataset tensors ``[N,.]``, ``index`` the bootstrap table ``[E,train_length]``
        (int64, device).  Returns ``(ensemble_loss [nsteps,E], loss_terms [nsteps,3])`` device
        tensors, no synchronisation."""
        model, L = self.model, _lib.lib()
        dev = model.device
        self._state_to(dev)
        e0, e1 = self.members
        E = model.ensemble_size
        el = torch.zeros(max(nsteps, 1), E, device=dev)
        terms = torch.zeros(max(nsteps, 1), 3, device=dev)
        if nsteps == 0 or e0 == e1:
            return el, terms
        with torch.cuda.device(dev):
            m = model._c_model(None)
            need = L.mb_pe_train_steps_workspace_bytes(ctypes.byref(m), batch_size)
            if need == 0:
                raise RuntimeError("the fused training kernel does not cover this model shape")
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
                self._graph = None
            cfg = model.train_cfg(self.ignore_terminal_state, self.lr, self.betas, self.eps)
            assert index.dtype == torch.int64 and index.is_contiguous() and index.shape[0] == E
            _lib.check(L.mb_pe_train_steps(
                ctypes.byref(m), ctypes.byref(cfg), _lib.ptr(model.module.flat.data), _lib.ptr(self.exp_avg),
                _lib.ptr(self.exp_avg_sq), self.num_train_steps, _lib.ptr(data["observations"]),
                _lib.ptr(data["actions"]), _lib.ptr(data["deltas"]), _lib.ptr(data["rewards"]),
                _lib.ptr(data["terminals"]), _lib.ptr(index), index.shape[1], train_length, batch_size,
                first_step, nsteps, e0, e1, _lib.ptr(el), _lib.ptr(terms), _lib.ptr(self._ws),
                self._ws.numel(), _lib.stream(dev)))
        self.num_train_steps += nsteps
        model.module.mark_weights_changed()
        return el, terms

    def train_from_torch_batch(self, batch):
        """pe_model_trainer.py:85-106 (same diagnostics keys; synchronises like the reference)."""
        ensemble_loss, terms = self.train_step_async(batch)
        return self._train_diagnostics(ensemble_loss, terms)

    def _train_diagnostics(self, ensemble_loss, terms):
        el = ensemble_loss.cpu().numpy()
        diagnostics = OrderedDict()
        e0, e1 = self.members
        for i in range(e0, e1):
            diagnostics["mt/net_%d/train_loss" % i] = el[i]
        diagnostics["train_loss"] = float(terms.sum().item())
        return diagnostics

    def eval_with_dataset(self, dataset):
        """pe_model_trainer.py:108-121: whole hold-out set, shared by all members."""
        dev = self.model.device
        d = {k: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v), device=dev).float())
             for k, v in dataset.items() if k in _FIELDS}
        loss, ensemble_loss = self.model.compute_loss(d["observations"], d["actions"], d["deltas"],
                                                      d["rewards"], d["terminals"])
        el = ensemble_loss.cpu().numpy()
        diagnostics = OrderedDict()
        for i, l in enumerate(el):
            diagnostics["mt/net_%d/eval_loss" % i] = l
        diagnostics["eval_loss"] = loss.item()
        return diagnostics, list(el)

    # ---- dataset handling ------------------------------------------------------------------------
    def set_mean_std(self, mean_std_dict=None):
        """pe_model_trainer.py:123-136."""
        model = self.model
        if mean_std_dict is None:
            mean_std_dict = self.mean_std_dict
        else:
            self.mean_std_dict = mean_std_dict
        model.obs_processor.set_mean_std_np(*mean_std_dict["observations"])
        model.action_processor.set_mean_std_np(*mean_std_dict["actions"])
        model.delta_processor.set_mean_std_np(*mean_std_dict["deltas"])

    def train_with_pool(self, pool):
        """pe_model_trainer.py:138-170."""
        max_step = self.max_step_per_train
        getter = getattr(pool, "get_unprocessed_data_torch", pool.get_unprocessed_data)     # device pool
        temp = getter("compute_delta", ["observations", "next_observations"])
        deltas = temp["next_observations"] - temp["observations"]
        pool.update_single_extra_field("deltas", deltas)
        pool.update_process_flag("compute_delta", len(deltas))
        self.set_mean_std(pool.get_mean_std())
        max_step = self.init_model_train_step if self.num_train_steps == 0 else int(max_step)
        if self.load_dir is not None:
            if not self.model.load(self.load_dir) and logger is not None:
                logger.log("Can not load a model from %s" % self.load_dir)
        results = self.sufficiently_train(
            pool, max_step, valid_ratio=self.valid_ratio, max_valid=self.max_valid,
            resample=self.resample, batch_size=self.batch_size, report_freq=self.report_freq,
            max_not_improve=self.max_not_improve, load_best=self.load_best, silent=self.silent,
            load_dataset_dir=self.load_dataset_dir)
        if self.save_dir is not None:
            os.makedirs(self.save_dir, exist_ok=True)
            self.model.save(self.save_dir)
        return results

    def split_dataset(self, pool, valid_ratio, max_valid, resample):
        """pe_model_trainer.py:172-182 (permutation, then the bootstrap index table)."""
        E = self.model.ensemble_size
        if hasattr(pool, "get_data_torch"):
            # device-resident pool: same permutation draw (pools/utils.py:69), applied on the device
            dataset = pool.get_data_torch()
            n_sample = len(next(iter(dataset.values())))
            valid_length = int(np.ceil(min(max_valid, valid_ratio * n_sample)))
            train_length = n_sample - valid_length
            indices = torch.from_numpy(np.random.permutation(n_sample)).to(self.model.device)
            tr_i, va_i = indices[:train_length], indices[n_sample - valid_length:]
            self.train_dataset = {k: v.index_select(0, tr_i) for k, v in dataset.items()}
            self.eval_dataset = {k: v.index_select(0, va_i) for k, v in dataset.items()}
            self.train_length, self.eval_length = train_length, valid_length
        else:
            dataset = pool.get_data()
            (self.train_dataset, self.eval_dataset,
             self.train_length, self.eval_length) = split_dataset(dataset, valid_ratio, max_valid)
        if resample:
            self.ensemble_index = np.random.randint(self.train_length, size=(E, self.train_length))
        else:
            self.ensemble_index = np.tile(np.arange(self.train_length)[None], (E, 1))

    def save_dataset(self, save_dir=None):
        save_dic = {k: getattr(self, k) for k in ("train_dataset", "eval_dataset", "train_length",
                                                  "eval_length", "ensemble_index", "mean_std_dict")}
        np.save(osp.join(save_dir or snapshot_dir(), "pe_model_dataset.npy"), save_dic,
                allow_pickle=True)

    def load_dataset(self, load_dir=None):
        path = osp.join(load_dir or snapshot_dir(), "pe_model_dataset.npy")
        self.__dict__.update(np.load(path, allow_pickle=True).item())

    # ---- device-resident early-stopping snapshots -------------------------------------------------
    def _snapshot_member(self, i):
        mod = self.model.module
        if self.snapshot_to_disk:
            self.model.save(net_id=i)
        if self._best is None or self._best.shape != mod.flat.shape or self._best.device != mod.flat.device:
            self._best = mod.flat.data.clone()
        for a, b in mod.member_slices(i, i + 1):
            self._best[a:b].copy_(mod.flat.data[a:b])

    def _restore_member(self, i):
        """``model.load(net_id=i)`` (pe_model_trainer.py:282-284).  Like the reference's
        substring match 'net{i}' over state-dict keys, the log-std bounds are restored too."""
        mod = self.model.module
        for a, b in mod.member_slices(i, i + 1):
            mod.flat.data[a:b].copy_(self._best[a:b])
        mod.mark_weights_changed()

    def sufficiently_train(self, pool=None, max_step=2000, valid_ratio=0.2, max_valid=5000,
                           resample=True, batch_size=256, report_freq=100, max_not_improve=20,
                           load_best=True, silent=False, load_dataset_dir=None):
        """pe_model_trainer.py:212-293.  Same schedule as the reference's ``while True: for j in
        range(0, train_length, batch_size)`` loop -- step ``t`` trains on bootstrap columns
        ``[k*batch_size, (k+1)*batch_size)`` with ``k = t mod ceil(train_length / batch_size)`` and the
        hold-out evaluation runs whenever ``t % report_freq == 0`` -- but the steps between two
        evaluations go out as ONE persistent launch.

        Member-sharded (``members`` a strict slice, ``torch.distributed`` initialised): every rank
        draws the same split / bootstrap table (same NumPy stream) and trains its own members; there is
        no collective in the training steps.  The reference keeps training EVERY member until all of
        them have stalled (pe_model_trainer.py:259-263), so the stop decision is global: one 4-byte
        all-reduce per evaluation point.  Afterwards the trained slices are all-gathered, so the final
        evaluation, ``remember_loss`` and therefore ``elite_indices`` are identical on every rank -- and
        identical to an unsharded run (members are independent).  An empty slice (E=7 over 8 ranks)
        trains nothing and still takes part in the collectives."""
        from . import dist as mbdist
        model = self.model
        dev = model.device
        if pool is not None:
            self.split_dataset(pool, valid_ratio, max_valid, resample)
        else:
            assert load_dataset_dir is not None
            self.load_dataset(load_dataset_dir)
            self.set_mean_std()

        def _dev(v):
            return (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(dev).float().contiguous()
        train = {k: _dev(v) for k, v in self.train_dataset.items() if k in _FIELDS}
        evald = {k: _dev(v) for k, v in self.eval_dataset.items() if k in _FIELDS}
        index = torch.as_tensor(self.ensemble_index).to(device=dev, dtype=torch.int64).contiguous()
        train_length = self.train_length
        e0, e1 = self.members
        fused = self.fused_steps_supported() and not self.use_cuda_graph
        total_train_step = 0
        eval_stat, eval_loss, last = OrderedDict(), None, None
        continue_training = True
        sharded = (e0, e1) != (0, model.ensemble_size) and mbdist.world()[1] > 1
        while True:
            if total_train_step % report_freq == 0:
                eval_stat, eval_loss = self.eval_with_dataset(evald)
                if total_train_step == 0:
                    min_loss = eval_loss
                    not_improve = []
                    for i, _ in enumerate(eval_loss):
                        self._snapshot_member(i)
                        not_improve.append(0)
                else:
                    for i, l in enumerate(eval_loss):
                        if l < min_loss[i]:
                            min_loss[i] = l
                            not_improve[i] = 0
                            self._snapshot_member(i)
                        else:
                            not_improve[i] += 1
                continue_training = False
                for i, n in enumerate(not_improve):
                    eval_stat["mt/net_%d/not_improve" % i] = n
                    if e0 <= i < e1 and (max_not_improve < 0 or n < max_not_improve):
                        continue_training = True
                if sharded:                                   # any member anywhere still improving?
                    flag = torch.tensor([float(continue_training)], device=dev)
                    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
                    continue_training = flag.item() > 0
            if total_train_step >= max_step:
                continue_training = False
            if not continue_training:
                break
            # steps until the next evaluation (or the step budget): one launch
            n = min(report_freq - total_train_step % report_freq, int(max_step) - total_train_step)
            if fused:
                el, terms = self.train_steps_async(train, index, train_length, batch_size, total_train_step, n)
                last = (el[n - 1], terms[n - 1])
            else:
                spe = (train_length + batch_size - 1) // batch_size
                for t in range(total_train_step, total_train_step + n):
                    j = (t % spe) * batch_size
                    ind = index[:, j:j + batch_size]
                    last = self.train_step_async({k: v[ind] for k, v in train.items()})
            total_train_step += n
        if load_best and e1 > e0 and self._best is not None:
            for i in range(e0, e1):
                self._restore_member(i)
        if sharded:
            parts = [None] * mbdist.world()[1]
            torch.distributed.all_gather_object(parts, (e0, e1))
            mbdist.all_gather_members(model.module, parts)   # every rank now holds every trained member
        eval_stat, eval_loss = self.eval_with_dataset(evald)
        model.remember_loss(eval_loss)
        if last is not None:
            self.statistics.update(self._train_diagnostics(*last))
        self.statistics.update(eval_stat)
        self.statistics["train_step"] = total_train_step
        self.min_loss = eval_loss
        return self.statistics

    def get_diagnostics(self):
        return self.statistics

    def end_epoch(self, epoch):
        pass

    @property
    def device(self):
        return self.model.device

    @property
    def networks(self):
        return [self.model]

    def get_snapshot(self):
        return dict(model=self.model)
