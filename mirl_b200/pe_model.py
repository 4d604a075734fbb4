"""B200PEModel -- drop-in for ``PEModel`` (mirl/agents/mbpo/pe_model.py:191-323 over
``Model``, mirl/agents/mbpo/base_model.py:7-76) whose ``step`` / ``compute_loss`` run as fused
sm_100a CUDA kernels through the C ABI in ``include/mirl_b200.h``.

Same constructor arguments, attributes (``ensemble_size``, ``elite_indices``, ``rank``,
``obs_processor`` ...), return values and state-dict key names as the reference.  Two optional
keyword arguments are added to ``step`` for reproducibility: ``indices`` and ``noise``; when
omitted they are drawn exactly like the reference draws them (``np.random.choice`` from the
global NumPy generator, pe_model.py:169, then ``torch.randn`` from the global torch generator of
the tensor's device, pe_model.py:176), in the same order.
"""
import collections
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .compat import Component, short_env_name, snapshot_dir
from .ensemble import PackedEnsembleMLP


class Normalizer(nn.Module):
    """mirl/processors/normalizer.py:6-28 (frozen mean/std, epsilon 1e-6)."""

    def __init__(self, shape, epsilon=1e-6):
        super().__init__()
        if abs(epsilon - 1e-6) > 1e-12:
            raise NotImplementedError("Normalizer epsilon is fixed at 1e-6 in the kernels")
        self.input_shape = self.output_shape = shape
        self.epsilon = epsilon
        self.mean = nn.Parameter(torch.zeros(shape), requires_grad=False)
        self.std = nn.Parameter(torch.ones(shape), requires_grad=False)

    def forward(self, x):
        return self.process(x)

    def process(self, x):
        return (x - self.mean) / (self.std + self.epsilon)

    def recover(self, x):
        return x * (self.std + self.epsilon) + self.mean

    def set_mean_std_np(self, mean, std):
        dev = self.mean.device
        self.mean.data = torch.as_tensor(np.asarray(mean), device=dev).float()
        self.std.data = torch.as_tensor(np.asarray(std), device=dev).float()

    def mean_std_np(self):
        return self.mean.data.cpu().numpy(), self.std.data.cpu().numpy()


def torch_done_function(kind):
    """``done_f`` attribute of the reference model (reward_done_functions/*.py), kept as a
    public callable because other components read ``model.done_f`` (omega_agent.py).  The
    rollout kernels do NOT call this -- the done mask is fused in CUDA."""
    def done_f(obs, action, next_obs, input_type="torch"):
        if input_type != "torch":
            raise NotImplementedError("numpy done functions are not part of the B200 path")
        if kind == "cheetah":
            return torch.zeros_like(next_obs[..., 0:1])
        if kind == "hopper":
            live = (torch.isfinite(next_obs).all(-1) & (torch.abs(next_obs[..., 1:]) < 100).all(-1)
                    & (next_obs[..., 0] > 0.7) & (torch.abs(next_obs[..., 1]) < 0.2))
            return (~live)[..., None].float()
        if kind == "walker2d":
            h, a = next_obs[..., 0:1], next_obs[..., 1:2]
            return (~((h > 0.8) & (h < 2.0) & (a > -1.0) & (a < 1.0))).float()
        if kind == "mb_ant":
            live = torch.isfinite(next_obs).all(-1) & (next_obs[..., 0] >= 0.2) & (next_obs[..., 0] <= 1)
            return (~live)[..., None].float()
        if kind == "mb_humanoid":
            q = next_obs[..., 0:1]
            return ((q < 1.0) | (q > 2.0)).float()
        if kind == "inverted_pendulum":
            live = (torch.isfinite(next_obs).all(-1) & (next_obs[..., 1] <= 0.2)
                    & (next_obs[..., 1] >= -0.2))
            return (~live)[..., None].float()
        raise NotImplementedError(kind)
    return done_f


class B200PEModel(Component, nn.Module):
    def __init__(self, env, normalize_obs=True, normalize_action=True, normalize_delta=True,
                 normalize_reward=False, known=["done"], ensemble_size=7, allow_to_use=None,
                 elite_number=5, model_name="PE_model", reward_coef=1, bound_reg_coef=2e-2,
                 weight_decay=[2.5e-5, 5e-5, 7.5e-5, 1e-5, 7.5e-5], kernel=None,
                 device_draw_min_rows=32768, **mlp_kwargs):
        nn.Module.__init__(self)
        self.env = env
        self.env_name = env.env_name
        known = known if isinstance(known, list) else [known]
        for item in known:
            assert item in ["reward", "done"]
        if "reward" in known or "done" not in known:
            # every reward_function raises NotImplementedError upstream and learn_done is
            # asserted off (pe_model.py:39): the path always learns the reward, knows done
            raise NotImplementedError("B200PEModel supports known=['done'] (learned reward, "
                                      "analytic done), like every MIRL config of this path")
        if not (normalize_obs and normalize_action and normalize_delta) or normalize_reward:
            raise NotImplementedError("B200PEModel fuses obs/action/delta normalisation and no "
                                      "reward normalisation (the reference defaults)")
        self.done_kind = short_env_name(self.env_name)
        if self.done_kind not in _lib.DONE_KIND:
            raise NotImplementedError("no done function for env %r" % self.env_name)
        self.reward_f, self.done_f = None, torch_done_function(self.done_kind)
        self.observation_shape = env.observation_space.shape
        self.action_shape = env.action_space.shape
        self.processed_obs_shape = self.observation_shape
        self.processed_action_shape = self.action_shape
        assert len(self.processed_obs_shape) == 1 and len(self.processed_action_shape) == 1
        self.need_to_learn = ["reward"]
        self.learn_reward, self.learn_done = True, False
        self.normalize_obs, self.normalize_action = normalize_obs, normalize_action
        self.normalize_delta, self.normalize_reward = normalize_delta, normalize_reward
        self.obs_processor = Normalizer(self.observation_shape)
        self.action_processor = Normalizer(self.action_shape)
        self.delta_processor = Normalizer(self.observation_shape)
        self.ensemble_size = ensemble_size
        self.elite_number = elite_number
        O, A = self.observation_shape[0], self.action_shape[0]
        self.obs_size, self.action_size = O, A
        hidden = mlp_kwargs.pop("hidden_layers")
        self.module = PackedEnsembleMLP(O + A, 2 * (O + 1), hidden, ensemble_size,
                                        module_name=model_name, probabilistic=True, **mlp_kwargs)
        if self.module.activation != "swish":
            raise NotImplementedError("the fused training kernel differentiates swish only")
        self.reward_coef = reward_coef
        self.bound_reg_coef = bound_reg_coef
        wd = weight_decay if isinstance(weight_decay, (list, tuple)) else [weight_decay]
        wd = list(wd) + [wd[-1]] * (self.module.num_fc - len(wd))     # to_list pads with the last
        assert len(wd) == self.module.num_fc                          # mlp.py:173-175
        self.weight_decay = wd
        if allow_to_use is None:
            allow_to_use = self.ensemble_size
        self.allow_to_use = allow_to_use
        self.elite_indices = None
        self.rank = list(np.arange(ensemble_size))
        self.loss = None
        self.kernel = kernel or os.environ.get("MIRL_B200_KERNEL", "auto")
        # batches at least this large draw their members on the device (same NumPy stream, see
        # draw_members); smaller ones call np.random.choice itself
        self.device_draw_min_rows = device_draw_min_rows
        self._norm_cache = (None, None)
        self._pack_cache = (None, None)
        self._ws = None
        self._draw_ws = None
        self._draw_stream = None
        self._draw_state = None
        self._draw_queue = collections.deque()
        self._draw_snaps = []
        self._draw_np_state = None
        self._draw_last_shape = None
        self.draw_lookahead = 2
        self._drew_on_device = False
        self.draw_overlap_sms = 1

    # ---- reference bookkeeping ---------------------------------------------------------------
    def remember_loss(self, loss):
        """pe_model.py:234-237."""
        self.loss = loss
        self.rank = np.argsort(loss[:self.allow_to_use])
        self.elite_indices = list(self.rank[:self.elite_number])

    def save(self, save_dir=None, net_id=None):
        self.module.save(save_dir or snapshot_dir(), net_id)

    def load(self, load_dir=None, net_id=None):
        return self.module.load(load_dir or snapshot_dir(), net_id)

    def get_snapshot(self):
        return self.module.get_snapshot()

    def get_diagnostics(self):
        return {}

    def reset(self):
        pass

    # ---- C-ABI plumbing ----------------------------------------------------------------------
    @property
    def device(self):
        return self.module.flat.device

    def _norm(self):
        ts = [self.obs_processor.mean, self.obs_processor.std, self.action_processor.mean,
              self.action_processor.std, self.delta_processor.mean, self.delta_processor.std]
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if self._norm_cache[0] != key:
            buf = torch.cat([t.detach().reshape(-1).float() for t in ts]).to(self.device).contiguous()
            self._norm_cache = (key, buf)
        return self._norm_cache[1]

    def _resolve_kernel(self):
        k = self.kernel
        if k == "auto":
            k = "tcgen05" if self._tc_supported() else "simt"
        if k not in _lib.KERNEL:
            raise ValueError("unknown kernel family %r" % (k,))
        return k

    def _tc_supported(self):
        L = _lib.lib()
        m = self._c_model(None)
        return L.mb_pe_tc_pack_bytes(ctypes.byref(m)) > 0

    def _c_model(self, pack):
        m = _lib.PeModel()
        m.mlp = self.module.desc
        m.obs_dim, m.act_dim = self.obs_size, self.action_size
        m.params = self.module.flat.data_ptr()
        m.norm = self._norm().data_ptr()
        m.tc_pack = pack.data_ptr() if pack is not None else None
        return m

    def _tc_pack(self):
        """bf16 hi/lo UMMA operand pack, rebuilt when the weights changed."""
        key = self.module.weights_key()
        if self._pack_cache[0] != key:
            L = _lib.lib()
            m = self._c_model(None)
            n = L.mb_pe_tc_pack_bytes(ctypes.byref(m))
            if n == 0:
                raise RuntimeError("tcgen05 rollout kernel does not support this model shape")
            pack = self._pack_cache[1]
            if pack is None or pack.numel() != n or pack.device != self.device:
                pack = torch.empty(n, dtype=torch.uint8, device=self.device)
            _lib.check(L.mb_pe_tc_pack(ctypes.byref(m), _lib.ptr(pack), _lib.stream()))
            self._pack_cache = (key, pack)
        return self._pack_cache[1]

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != self.device:
            self._ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _member_indices(self, B, indices):
        elite = self.elite_indices
        if elite is None:
            elite = list(np.arange(self.ensemble_size))               # pe_model.py:167-168
        self._drew_on_device = False
        if indices is None:
            if B >= self.device_draw_min_rows and len(elite) <= 64:
                self._drew_on_device = True
                return self.draw_members(B)
            indices = np.random.choice(elite, B)                      # pe_model.py:169
        if torch.is_tensor(indices):
            return indices.to(device=self.device, dtype=torch.uint8).contiguous()
        idx8 = np.ascontiguousarray(np.asarray(indices).astype(np.uint8))
        return torch.from_numpy(idx8).to(self.device, non_blocking=True)

    def _reserve_for_draw(self, L):
        """The rollout kernel is persistent (one CTA per SM) and the generator of the member draw
        is one CTA on a side stream: when this model draws the members itself, the rollout leaves
        ``draw_overlap_sms`` SMs free so that the NEXT call's draw runs under this call's rollout
        instead of behind it (measured: rollout + draw 1.57 -> 1.05 ms per 1e6 rows)."""
        L.mb_set_reserved_sms(self.draw_overlap_sms if self._drew_on_device else 0)

    def _draw_launch(self, start, el, n, out):
        """Queue one member draw on the draw stream.  The generator state lives in a PINNED host
        block the kernels read and advance in place (mapped memory, same pointer on both sides):
        no copy-engine transfer, so neither direction queues behind bulk H2D / D2H copies the
        caller has in flight, and a draw queued behind another continues the stream with no host
        round trip.  ``start`` = (key, pos) re-seeds the block from the host first (the draw
        stream must be idle then); None chains on whatever the previous draw left.  Returns
        (event, snapshot): the snapshot block receives this draw's final state."""
        L = _lib.lib()
        side = self._draw_stream
        if start is not None:
            host = self._draw_state.numpy().view(np.uint32)
            host[:624], host[624] = start
            host[625:] = 0
        snap = self._draw_snaps.pop() if self._draw_snaps else torch.zeros(627, dtype=torch.int32).pin_memory()
        nbytes = L.mb_pe_draw_members_workspace_bytes(n, len(el))
        with torch.cuda.stream(side):
            if self._draw_ws is None or self._draw_ws.numel() < nbytes or self._draw_ws.device != out.device:
                self._draw_ws = torch.empty(nbytes + 4096, dtype=torch.uint8, device=out.device)
            _lib.check(L.mb_pe_draw_members_chained(
                ctypes.c_void_p(self._draw_state.data_ptr()), ctypes.c_void_p(snap.data_ptr()), el, len(el), n,
                _lib.ptr(out), _lib.ptr(self._draw_ws), self._draw_ws.numel(), ctypes.c_void_p(side.cuda_stream)))
            ev = torch.cuda.Event()
            ev.record(side)
        return ev, snap

    def _draw_result(self, ev, snap, el):
        """Wait for a queued draw; -> (key, pos, produced) from its snapshot block (recycled)."""
        ev.synchronize()
        host = snap.numpy().view(np.uint32)
        key, pos = (host[:624].copy(), int(host[624])) if len(el) > 1 else (None, None)
        produced = int(host[625]) | (int(host[626]) << 32)
        self._draw_snaps.append(snap)
        return key, pos, produced

    @_lib.on_device
    def draw_members(self, B):
        """``np.random.choice(elite, B)`` (pe_model.py:165-169) generated on the device from
        NumPy's own global MT19937 state: the returned uint8 indices are bit-identical to the
        host call and ``np.random`` is left in the state that call would leave it in.  Costs one
        read of a 627-word state block (an event wait) instead of a 10-25 ms host loop per
        million rows.

        The draw runs on its own stream.  When consecutive calls ask for the same (B, elites),
        the next ``draw_lookahead`` draws are queued ahead, each continuing the stream on the
        device where the previous one stopped.  A later call consumes the head of that queue only
        if ``np.random`` is still exactly in the state the previous call left (nobody else drew)
        and the shape is the same; otherwise the queue is dropped and the draw starts from the
        host state.  Either way the result is what a fresh draw gives -- but a queued one was
        generated under earlier rollouts, off the host's critical path (small batches are
        otherwise paced by the draw's launch-and-wait latency, ~0.3 ms per call)."""
        elite = self.elite_indices
        if elite is None:
            elite = list(np.arange(self.ensemble_size))
        dev = self.device
        st = np.random.get_state()
        assert st[0] == "MT19937"
        key, pos = st[1], int(st[2])
        el = bytes(int(e) for e in elite)
        main = torch.cuda.current_stream(dev)
        if self._draw_stream is None:
            self._draw_stream = torch.cuda.Stream(dev)
            self._draw_state = torch.zeros(627, dtype=torch.int32).pin_memory()
        side = self._draw_stream
        shape = (B, el, str(dev))
        q = self._draw_queue
        left = self._draw_np_state
        hit = bool(q) and q[0]["shape"] == shape and left is not None and left[1] == pos and \
            np.array_equal(left[0], key)
        if q and not hit:                          # drop the look-ahead; the state block must be idle
            for d in q:
                self._draw_result(d["event"], d["snap"], el)
            q.clear()
        out, done, ev = None, 0, None
        if hit:
            d = q.popleft()
            out, ev = d["out"], d["event"]
            key, pos, done = self._draw_result(ev, d["snap"], el)
            assert 0 < done <= B
            if done < B:                           # word budget ran out: the queued draws behind it are void
                for d in q:
                    self._draw_result(d["event"], d["snap"], el)
                q.clear()
        if out is None:
            with torch.cuda.stream(side):
                out = torch.empty(B, dtype=torch.uint8, device=dev)
        while done < B:      # one pass unless the word budget (mean + 12 sigma attempts) ran out
            ev, snap = self._draw_launch((key, pos), el, B - done, out[done:])
            k2, p2, produced = self._draw_result(ev, snap, el)
            if len(el) > 1:
                key, pos = k2, p2
            assert 0 < produced <= B - done
            done += produced
        main.wait_event(ev)
        out.record_stream(main)
        np.random.set_state((st[0], key, pos, st[3], st[4]))
        self._draw_np_state = (np.array(key, copy=True), pos)
        if shape == self._draw_last_shape and len(el) > 1:
            while len(q) < self.draw_lookahead:    # every queued draw chains on the block in place
                with torch.cuda.stream(side):
                    nxt = torch.empty(B, dtype=torch.uint8, device=dev)
                e2, s2 = self._draw_launch(None, el, B, nxt)
                q.append({"shape": shape, "out": nxt, "event": e2, "snap": s2})
        self._draw_last_shape = shape
        return out

    # ---- rollout -------------------------------------------------------------------------------
    @_lib.on_device
    def step(self, obs, action, return_distribution=False, indices=None, noise=None,
             penalty_type=None, penalty_coef=1.0):
        """``PEModel.step`` (pe_model.py:252-283): ``(next_obs [B,O], reward [B,1], done [B,1],
        info)``.  With ``return_distribution`` ``info`` holds the six per-elite tensors in rank
        order; ``penalty_type`` ('total_var' | 'max_var') additionally fuses MOPO's penalty
        (mopo_collector.py:42-53) and returns ``info['penalty']`` / ``info['penalized_reward']``.
        """
        assert obs.dim() == action.dim() == 2
        dev = self.device
        L = _lib.lib()
        obs, action = _lib.f32(obs, dev), _lib.f32(action, dev)
        B, O, D = obs.shape[0], self.obs_size, self.obs_size + 1
        idx = self._member_indices(B, indices)
        if noise is None:
            noise = torch.randn(B, D, device=dev)                     # pe_model.py:176
        noise = _lib.f32(noise, dev)
        assert noise.shape == (B, D) and idx.shape == (B,)
        next_obs = torch.empty(B, O, device=dev)
        reward = torch.empty(B, 1, device=dev)
        done = torch.empty(B, 1, device=dev)
        info = {}
        if B == 0:
            if penalty_type is not None:                 # a rank whose rows have all terminated
                info["penalty"] = torch.empty(0, 1, device=dev)
                info["penalized_reward"] = torch.empty(0, 1, device=dev)
            return next_obs, reward, done, info
        kind = _lib.DONE_KIND[self.done_kind]
        if not return_distribution and penalty_type is None:
            kname = self._resolve_kernel()
            pack = self._tc_pack() if kname == "tcgen05" else None
            m = self._c_model(pack)
            ws = self._workspace(L.mb_pe_rollout_workspace_bytes(ctypes.byref(m), B))
            self._reserve_for_draw(L)
            _lib.check(L.mb_pe_rollout_step(
                ctypes.byref(m), _lib.KERNEL[kname], _lib.ptr(obs), _lib.ptr(action), _lib.ptr(idx),
                _lib.ptr(noise), B, kind, _lib.ptr(next_obs), _lib.ptr(reward), _lib.ptr(done),
                _lib.ptr(ws), ws.numel(), _lib.stream()))
            L.mb_set_reserved_sms(0)
            return next_obs, reward, done, info
        elites = self.elite_indices
        if elites is None:
            raise TypeError("return_distribution before remember_loss(): elite_indices is None "
                            "(the reference fails at pe_model.py:260 the same way)")
        n = len(elites)
        el = (ctypes.c_int32 * n)(*[int(e) for e in elites])
        m = self._c_model(None)
        out = {k: torch.empty(n, B, O, device=dev) for k in
               ("ensemble_delta_mean", "ensemble_delta_std", "ensemble_next_obs_mean")}
        out["ensemble_reward_mean"] = torch.empty(n, B, 1, device=dev)
        out["ensemble_reward_std"] = torch.empty(n, B, 1, device=dev)
        pen = pr = None
        if penalty_type is not None:
            pen = torch.empty(B, 1, device=dev)
            pr = torch.empty(B, 1, device=dev)
        args = (_lib.ptr(obs), _lib.ptr(action), _lib.ptr(idx), _lib.ptr(noise), B, kind,
                el, n, _lib.ptr(next_obs), _lib.ptr(reward), _lib.ptr(done),
                _lib.ptr(out["ensemble_delta_mean"]), _lib.ptr(out["ensemble_delta_std"]),
                _lib.ptr(out["ensemble_next_obs_mean"]), _lib.ptr(out["ensemble_reward_mean"]),
                _lib.ptr(out["ensemble_reward_std"]), _lib.PENALTY[penalty_type], float(penalty_coef),
                _lib.ptr(pen), _lib.ptr(pr))
        if self._resolve_kernel() == "tcgen05":
            # all-elite forward on the tensor cores, then one elementwise kernel
            m = self._c_model(self._tc_pack())
            ws = self._workspace(L.mb_pe_dist_workspace_bytes(ctypes.byref(m), B, n))
            _lib.check(L.mb_pe_rollout_step_dist_tc(ctypes.byref(m), *args, _lib.ptr(ws), ws.numel(), _lib.stream()))
        else:
            _lib.check(L.mb_pe_rollout_step_dist(ctypes.byref(m), *args, _lib.stream()))
        # key order and aliasing follow pe_model.py:255-282
        info["ensemble_delta_mean"] = out["ensemble_delta_mean"]
        info["ensemble_next_obs_mean"] = out["ensemble_next_obs_mean"]
        info["ensemble_delta_std"] = out["ensemble_delta_std"]
        info["ensemble_next_obs_std"] = out["ensemble_delta_std"]
        info["ensemble_reward_mean"] = out["ensemble_reward_mean"]
        info["ensemble_reward_std"] = out["ensemble_reward_std"]
        if penalty_type is not None:
            info["penalty"], info["penalized_reward"] = pen, pr
        return next_obs, reward, done, info

    def row_width(self):
        """Floats per pool row ``[obs | act | next_obs | reward | done]``."""
        return 2 * self.obs_size + self.action_size + 2

    def row_stride(self):
        """Row width rounded up to 16 bytes: the stride a device pool needs for bucket-order
        (bulk-copy) output."""
        return (self.row_width() + 3) // 4 * 4

    @_lib.on_device
    def step_rows(self, obs, action, dest, row_offset=0, indices=None, noise=None, ordered=True):
        """``PEModel.step`` fused with the collector's pool-row assembly (mbpo_collector.py:69-83):
        every transition is written as a row ``[obs | act | next_obs | reward | done | 0-padding]``
        into rows ``[row_offset, row_offset + B)`` of every buffer in ``dest``.

        ``dest`` is one ``[rows, stride]`` float32 CUDA tensor (``row_width() <= stride <=
        row_width() + 8``), or a list of up to 8 destinations mixing such tensors and raw device
        pointers (ints) -- peer-mapped pools of the other GPUs, which the kernel fills over NVLink
        itself (see ``dist.PeerPools``).  ``ordered=True`` puts batch row ``b`` at
        ``row_offset + b``; ``ordered=False`` writes in bucket order (grouped by drawn member):
        a pool is a bag of samples, and whole 128-row tiles then leave as one bulk copy per
        destination -- the mode to use for replicated pools (needs ``stride % 4 == 0``).
        ``obs`` / ``action`` may be column slices of a previous step's rows (row stride > width),
        so a horizon loop chains on the ``next_obs`` columns without a copy.  Returns the local
        destination's ``(next_obs, reward, done)`` views when ``dest[0]`` is a tensor."""
        dev = self.device
        L = _lib.lib()
        B, O, A, D = obs.shape[0], self.obs_size, self.action_size, self.obs_size + 1
        W = self.row_width()
        for t, w in ((obs, O), (action, A)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == w
                    and (B == 0 or t.stride(1) == 1)):
                raise ValueError("step_rows: obs/action must be float32 CUDA [B, width] with unit column stride")
        dests = list(dest) if isinstance(dest, (list, tuple)) else [dest]
        if not 1 <= len(dests) <= 8:
            raise ValueError("step_rows: 1..8 destinations")
        if not torch.is_tensor(dests[0]):
            raise ValueError("step_rows: the first destination must be a tensor (it defines the row stride)")
        stride = dests[0].shape[1] if dests[0].dim() == 2 else -1
        ptrs = []
        for d in dests:
            if torch.is_tensor(d):
                if not (d.is_cuda and d.dtype == torch.float32 and d.dim() == 2 and d.shape[1] == stride
                        and W <= stride <= W + 8 and d.is_contiguous() and d.shape[0] >= row_offset + B):
                    raise ValueError(f"step_rows: destination must be contiguous float32 CUDA "
                                     f"[>= {row_offset + B}, {W}..{W + 8}]")
                ptrs.append(d.data_ptr())
            else:
                ptrs.append(int(d))
        idx = self._member_indices(B, indices)
        if noise is None:
            noise = torch.randn(B, D, device=dev)                     # pe_model.py:176
        noise = _lib.f32(noise, dev)
        assert noise.shape == (B, D) and idx.shape == (B,)
        if B > 0:
            kname = self._resolve_kernel()
            pack = self._tc_pack() if kname == "tcgen05" else None
            m = self._c_model(pack)
            ws = self._workspace(L.mb_pe_rollout_workspace_bytes(ctypes.byref(m), B))
            arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
            self._reserve_for_draw(L)
            _lib.check(L.mb_pe_rollout_step_rows(
                ctypes.byref(m), _lib.KERNEL[kname], obs.data_ptr(), obs.stride(0), action.data_ptr(),
                action.stride(0), _lib.ptr(idx), _lib.ptr(noise), B, _lib.DONE_KIND[self.done_kind],
                arr, len(ptrs), int(stride), int(row_offset), int(bool(ordered)), _lib.ptr(ws), ws.numel(),
                _lib.stream()))
            L.mb_set_reserved_sms(0)
        rows = dests[0][row_offset:row_offset + B]
        return rows[:, O + A:2 * O + A], rows[:, 2 * O + A:2 * O + A + 1], rows[:, 2 * O + A + 1:W]

    def step_np(self, obs, action, **kwargs):
        """base_model.py:78-86."""
        dev = self.device
        o = torch.as_tensor(obs, device=dev).float()
        a = torch.as_tensor(action, device=dev).float()
        next_obs, reward, done, info = self.step(o, a, **kwargs)
        return (next_obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy(),
                {k: v.cpu().numpy() for k, v in info.items()})

    # ---- distribution / loss ---------------------------------------------------------------------
    @_lib.on_device
    def predict_distribution(self, obs, action, normalize=True):
        """Per-member Gaussian of the NORMALISED delta/reward: ``mean, logstd`` ``[E,B,O+1]``
        (MLP forward + ``_get_mean_logstd``, pe_model.py:110-151).  Inputs ``[B,.]`` are shared
        by all members, ``[E,B,.]`` are per member (linear.py:144-149)."""
        dev = self.device
        obs, action = _lib.f32(obs, dev), _lib.f32(action, dev)
        per_member = obs.dim() == 3
        B, E, D = obs.shape[-2], self.ensemble_size, self.obs_size + 1
        if per_member:
            assert obs.shape[0] == E and action.shape[0] == E
        if normalize and self._resolve_kernel() == "tcgen05":
            # tensor-core forward: [E][B][mean D | log-std D]; the two halves are returned as views
            m = self._c_model(self._tc_pack())
            both = torch.empty(E, B, 2 * D, device=dev)
            mem = (ctypes.c_int32 * E)(*range(E))
            if B > 0:
                _lib.check(_lib.lib().mb_pe_forward_tc(ctypes.byref(m), _lib.ptr(obs), _lib.ptr(action),
                                                       int(per_member), B, mem, E, _lib.ptr(both), _lib.stream()))
            return both[..., :D], both[..., D:]
        mean = torch.empty(E, B, D, device=dev)
        logstd = torch.empty(E, B, D, device=dev)
        m = self._c_model(None)
        _lib.check(_lib.lib().mb_pe_forward(ctypes.byref(m), _lib.ptr(obs), _lib.ptr(action),
                                            int(per_member), int(normalize), B, _lib.ptr(mean),
                                            _lib.ptr(logstd), _lib.stream()))
        return mean, logstd

    def train_cfg(self, ignore_terminal_state=False, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        c = _lib.TrainCfg()
        for l, wd in enumerate(self.weight_decay):
            c.weight_decay[l] = float(wd)
        c.bound_reg_coef = float(self.bound_reg_coef)
        c.reward_coef = float(self.reward_coef)
        c.ignore_terminal_state = int(bool(ignore_terminal_state))
        c.lr, c.beta1, c.beta2, c.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        return c

    def regularizers(self):
        """(l2, bound) terms of the loss (pe_model.py:80-81): small reductions over the blob."""
        mod = self.module
        l2 = mod.get_weight_decay(self.weight_decay)
        bound = self.bound_reg_coef * (mod.logstd_bound(0).sum() - mod.logstd_bound(1).sum())
        return l2, bound

    @_lib.on_device
    def compute_loss(self, obs, action, delta, reward=None, done=None, ignore_terminal_state=True,
                     n_step=0, log=False):
        """``PEModel.compute_loss`` (pe_model.py:285-307): ``(loss, ensemble_loss [E])`` for a
        shared ``[B,.]`` or per-member ``[E,B,.]`` batch.  Forward-only (evaluation): the
        gradient path is the fused ``B200PEModelTrainer`` step, so ``loss`` carries no graph."""
        assert reward is not None and self.reward_coef > 0            # pe_model.py:127
        dev = self.device
        obs, action, delta, reward = (_lib.f32(t, dev) for t in (obs, action, delta, reward))
        if ignore_terminal_state:
            assert done is not None                                   # pe_model.py:85
        done = _lib.f32(done, dev) if done is not None else None
        per_member = obs.dim() == 3
        B = obs.shape[-2]
        ensemble_loss = torch.empty(self.ensemble_size, device=dev)
        cfg = self.train_cfg(ignore_terminal_state)
        m = self._c_model(None)
        _lib.check(_lib.lib().mb_pe_loss(ctypes.byref(m), ctypes.byref(cfg), _lib.ptr(obs),
                                         _lib.ptr(action), _lib.ptr(delta), _lib.ptr(reward),
                                         _lib.ptr(done), int(per_member), B, _lib.ptr(ensemble_loss),
                                         None, _lib.stream()))
        l2, bound = self.regularizers()
        return ensemble_loss.sum() + l2 + bound, ensemble_loss

    def log_prob_np(self, *args, **kwargs):
        raise NotImplementedError
