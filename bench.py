#!/usr/bin/env python
"""bench.py -- MBPO ensemble rollout throughput (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one model rollout of HORIZON steps over ROWS start states per GPU (HalfCheetah-shaped
synthetic data: obs 17, act 6, 7 members / 5 elites, 4x200 swish), i.e. ROWS*HORIZON transitions
per GPU per step; actions are pre-generated U(-1,1) (the reference's own
``COMPOCollector(random_sample=True)`` mode), the done function is cheetah's.  Work per GPU is
fixed as N grows ("scaling": "weak"); at N > 1 every rank also all-gathers the transitions it
generated into a replicated pool buffer (NCCL), inside the timed region.

``value``   transitions/s, inputs resident in HBM, CUDA-event timed, max over ranks.
``e2e``     same metric through the public API (``B200PEModel.step``) from pinned HOST buffers:
            H2D of start states / actions / member indices and D2H of the generated
            transitions are inside the timed region, as are the NumPy member draw and the
            torch noise draw the reference also performs.
``roofline`` the rollout kernel alone (events inside the C ABI), algorithmic FLOPs / launch.
``cpu_baseline`` the CPU oracle (restatement of the reference's op sequence: all E members for
            every row) on this box's host cores, bounded sample.
``--impl reference`` times only that CPU arm (the reference itself is Python and cannot travel
            to the GPU box; the oracle is pinned to it by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = int(os.environ.get("MB_BENCH_ROWS", 1_000_000))       # start states per GPU per step
HORIZON = int(os.environ.get("MB_BENCH_HORIZON", 5))
CPU_SAMPLE_ROWS = int(os.environ.get("MB_BENCH_CPU_ROWS", 50_000))
FLOP_PER_TRANSITION = 263_600     # one member: 2*(23*200 + 3*200*200 + 200*36), SURVEY.md 8(d)
BYTES_PER_TRANSITION = 241        # obs 68 + act 24 + noise 72 + idx 1 in; next_obs 68 + r 4 + d 4 out
METRIC = "mbpo_rollout_transitions_per_s"
LOSS_VEC7 = [0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tensor": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"], "src": "measured (sustained)"}
    return {"tensor": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device):
    from tests.golden import cases
    from tests.helpers_gpu import make_pe_model
    m, p = make_pe_model(cases.MBPO, 11, "cheetah", device=device)
    m.remember_loss(LOSS_VEC7)
    return m, p


def cpu_arm(steps, warmup, threads=None):
    """Reference CPU path (oracle restatement, kind 'port'): PEModel.step over a bounded sample."""
    from oracle import pe_oracle as po
    from tests.golden import cases
    from tests.util import T, to_torch
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    p = to_torch(cases.pe_params(11, cases.MBPO))
    _, elites = po.rank_elites(LOSS_VEC7, 7, 5)
    B = CPU_SAMPLE_ROWS
    o, a = cases.rollout_inputs(3, B)
    o, a = T(o), T(a)
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            idx = torch.from_numpy(np.random.choice(elites, B))      # pe_model.py:169
            eps = torch.randn(B, 18)                                 # pe_model.py:176
            po.rollout_step(p, o, a, idx, eps, "cheetah")
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": B / sec, "unit": "transitions/s", "cores": threads, "kind": "port",
            "sample": "%d steps of PEModel.step restatement (all 7 members, fp32 torch CPU) over "
                      "%d states, %.2f s/step" % (steps, B, sec)}, sec


def secondary_metrics(dev, world, rank, timed):
    """Model-train samples/s (members sharded over ranks, no collective in the step) and
    REDQ / TQC critic-target throughput at batch 256 (replicas), device-resident inputs."""
    import mirl_b200
    from mirl_b200.dist import member_partition
    from tests.golden import cases
    from tests.helpers_gpu import FakeEnv, make_critic, make_pe_model
    from tests.util import T
    E, B = 7, 256
    m, _ = make_pe_model(cases.MBPO, 11, device=dev)
    e0, e1 = member_partition(E, world)[rank]
    out = {}
    if e1 > e0:
        tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, tb_log_freq=-1, silent=True, members=(e0, e1))
        bt = {k: T(v, dev) for k, v in cases.train_batch(600, E, B).items()}
        ms = timed(lambda: tr.train_step_async(bt), 200, 20) / 200
    else:
        ms = timed(lambda: None, 200, 20) / 200
    out["model_train_samples_per_s"] = E * B / (ms * 1e-3)
    out["model_train_ms_per_step"] = ms
    # forward + dX + dW = 3 GEMM passes of 2*sum(K_l*N_l) flops per sample (4x200 MLP, 23 -> 36)
    mlp_macs = 23 * 200 + 3 * 200 * 200 + 200 * 36
    out["model_train_tflops"] = 3 * 2 * mlp_macs * E * B / (ms * 1e-3) / 1e12
    out["model_train_roofline_frac"] = out["model_train_tflops"] / peaks()["tensor"]
    out["model_train_config"] = "7 members x batch 256 (bootstrap batches), fused fwd+NLL+bwd+Adam, members sharded over %d GPU(s)" % world
    q, _ = make_critic("redq", 71, device=dev)
    t, _ = make_critic("tqc", 81, device=dev)
    x = cases.critic_inputs(72, B)
    o, a = T(x["obs"], dev), T(x["act"], dev)
    r, d, lp = T(x["rewards"], dev), T(x["terminals"], dev), T(x["log_prob"], dev)
    ms = timed(lambda: q.q_target(o, a, r, d, lp, alpha=0.2, sample_number=2, batchwise_sample=True), 200, 20) / 200
    out["redq_targets_per_s"] = world * B / (ms * 1e-3)
    out["redq_tflops"] = 2 * 2 * (23 * 256 + 256 * 256 + 256) * B / (ms * 1e-3) / 1e12      # 2 drawn members
    ms = timed(lambda: t.atoms_target(o, a, r, d, lp, alpha=0.2, percentage_drop=8), 200, 20) / 200
    out["tqc_targets_per_s"] = world * B / (ms * 1e-3)
    out["tqc_tflops"] = 5 * 2 * (23 * 512 + 2 * 512 * 512 + 512 * 25) * B / (ms * 1e-3) / 1e12
    # configs/offline_mopo.json: penalised rollout step (all elites per row + max-var penalty)
    mm, _ = make_pe_model(cases.MBPO, 11, "hopper", device=dev)
    mm.remember_loss(LOSS_VEC7)
    Bm = 100_000
    gm = torch.Generator(device=dev).manual_seed(77)
    om = torch.randn(Bm, 17, device=dev, generator=gm)
    am = torch.rand(Bm, 6, device=dev, generator=gm) * 2 - 1
    em = torch.randn(Bm, 18, device=dev, generator=gm)
    im = torch.from_numpy(np.random.RandomState(5).choice(mm.elite_indices, Bm).astype(np.uint8)).to(dev)
    ms = timed(lambda: mm.step(om, am, indices=im, noise=em, return_distribution=True, penalty_type="max_var"), 20, 3) / 20
    out["mopo_penalised_transitions_per_s"] = world * Bm / (ms * 1e-3)
    out["mopo_config"] = "100k states, 5 elites evaluated per row (tcgen05 forward mode) + max_var penalty, per GPU replica"
    out["critic_config"] = "batch 256; REDQ 10x(2x256) random-2 min + TD; TQC 5x(3x512)x25 drop 10 + TD; replicas"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    base, sec = cpu_arm(steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "transitions/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "transitions/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "MBPO model rollout sweep point: %d start states/GPU x horizon %d, 7 members / "
                        "5 elites, 4x200 swish, obs17/act6, cheetah done" % (ROWS, HORIZON),
            "rows_per_gpu": ROWS, "horizon": HORIZON, "ensemble": 7, "elites": 5,
            "parallelism": "rows sharded over %d GPU(s); every transition row is stored into all %d pool replicas "
                           "by the rollout kernel itself (peer stores over NVLink), one 4-byte all-reduce per step" % (n, n),
            "l2_policy": "inputs larger than L2 (%.0f MB of obs/act/noise per step)" %
                         (ROWS * HORIZON * (24 + 72 + 1) / 1e6 + ROWS * 68 / 1e6)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from mirl_b200 import _lib
    L = _lib.lib()
    model, _ = build_model(dev)
    model.draw_lookahead = int(os.environ.get("MB_DRAW_AHEAD", "2"))               # A/B switches of the e2e leg
    model.draw_overlap_sms = int(os.environ.get("MB_DRAW_SMS", "1"))
    kernel = model._resolve_kernel()
    B, H, O, A, D = ROWS, HORIZON, 17, 6, 18
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    obs0 = torch.randn(B, O, device=dev, generator=gen)
    acts = torch.rand(H, B, A, device=dev, generator=gen) * 2 - 1
    noise = torch.randn(H, B, D, device=dev, generator=gen)
    rs = np.random.RandomState(99 + rank)
    idx_host = rs.choice(model.elite_indices, (H, B)).astype(np.uint8)
    idx = torch.from_numpy(idx_host).to(dev)
    # Model pool of generated transitions, rows [o | a | o' | r | d] (42 floats), replicated on every
    # rank: [H][world][B] rows.  The rollout kernel writes each row into every replica from its
    # epilogue (local HBM + peer-mapped NVLink stores), so there is no separate gather pass.
    from mirl_b200.dist import PeerPools
    pools = PeerPools(H * world * B, model.row_stride(), device=dev)     # 44-float rows (16-byte aligned)
    dests = pools.destinations()

    def device_step():
        o = obs0
        for h in range(H):
            # bucket order: a pool is a bag of samples, and a 128-row tile then leaves the SM as
            # one bulk copy per replica; the next step chains on the next_obs columns in place
            o, _, _ = model.step_rows(o, acts[h], dests, row_offset=(h * world + rank) * B,
                                      indices=idx[h], noise=noise[h], ordered=False)
        pools.fence()     # N>1: one stream-ordered 4-byte all-reduce = every replica is complete

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(device_step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B * H / (ms_per_step * 1e-3)

    # ---- end to end through the public API from pinned host buffers ------------------------------
    h_obs0 = obs0.cpu().pin_memory()
    h_acts = acts.cpu().pin_memory()
    np.random.seed(4321 + rank)
    torch.manual_seed(4321 + rank)

    # Public-API loop a user writes: start states and actions live in pinned HOST memory, the model
    # draws members (NumPy stream) and noise (torch) itself like PEModel.step, and every step's
    # [next_obs | reward | done] goes back to pinned host memory.  PCIe is full duplex, so the
    # batch is rolled in row chunks: the H2D copies run ahead on one copy stream in the order the
    # rollouts consume them, the D2H of every (chunk, step) runs on another under the next rollout
    # (two row buffers), and the link carries both directions at once.  Consecutive steps are
    # pipelined MB_E2E_DEPTH deep (default 2: double-buffered device inputs and host outputs; the
    # host takes ownership of step k's transitions before it issues step k+2; depth 1 = the
    # caller waits for every step's transitions before issuing the next step).
    W = 2 * O + A + 2
    n_chunks = max(1, int(os.environ.get("MB_E2E_CHUNKS", "2")))
    depth = max(1, min(2, int(os.environ.get("MB_E2E_DEPTH", "2"))))
    bounds = [B * c // n_chunks for c in range(n_chunks + 1)]
    Bc = max(bounds[c + 1] - bounds[c] for c in range(n_chunks))
    e2e_rows = torch.empty(2, Bc, W, device=dev)
    h_out = [torch.empty(H, B, O + 2).pin_memory() for _ in range(depth)]
    d_obs0 = [torch.empty(B, O, device=dev) for _ in range(depth)]
    d_acts = [torch.empty(H, B, A, device=dev) for _ in range(depth)]
    in_stream, copy_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    in_ready = [[[torch.cuda.Event() for _ in range(H)] for _ in range(n_chunks)] for _ in range(depth)]
    rolled = [torch.cuda.Event() for _ in range(depth)]        # the step's rollouts have read its device inputs
    landed = [torch.cuda.Event() for _ in range(depth)]        # the step's transitions are in host memory
    e2e_count = [0, 0]                                         # steps issued, row-buffer uses

    def e2e_step():
        main = torch.cuda.current_stream()
        step_no = e2e_count[0]
        p = step_no % depth
        e2e_count[0] += 1
        if step_no >= depth:
            landed[p].synchronize()                # the caller owns the transitions of step_no - depth
            in_stream.wait_event(rolled[p])        # ... and that step is done with this input set
        with torch.cuda.stream(in_stream):
            for c in range(n_chunks):
                lo, hi = bounds[c], bounds[c + 1]
                d_obs0[p][lo:hi].copy_(h_obs0[lo:hi], non_blocking=True)
                for h in range(H):
                    d_acts[p][h, lo:hi].copy_(h_acts[h, lo:hi], non_blocking=True)
                    in_ready[p][c][h].record(in_stream)
        for c in range(n_chunks):
            lo, hi = bounds[c], bounds[c + 1]
            o = d_obs0[p][lo:hi]
            for h in range(H):
                k = e2e_count[1]
                e2e_count[1] += 1
                buf = e2e_rows[k & 1, :hi - lo]
                main.wait_event(in_ready[p][c][h])
                main.wait_event(copy_done[k & 1])  # its previous contents have left for the host
                o, _, _ = model.step_rows(o, d_acts[p][h, lo:hi], buf)
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev)
                    h_out[p][h, lo:hi].copy_(buf[:, O + A:], non_blocking=True)
                    copy_done[k & 1].record(copy_stream)
        rolled[p].record(main)
        landed[p].record(copy_stream)
        if depth == 1:
            landed[p].synchronize()                # the caller owns the transitions now

    e2e_steps = max(2, min(args.steps, 5))
    e2e_ms = timed(e2e_step, e2e_steps, 3) / e2e_steps
    e2e = {"value": world * B * H / (e2e_ms * 1e-3), "unit": "transitions/s",
           "h2d_bytes_per_step": B * O * 4 + H * (B * A * 4 + n_chunks * 627 * 4),
           "d2h_bytes_per_step": H * B * (O + 2) * 4, "ms_per_step": e2e_ms, "row_chunks": n_chunks,
           "steps_in_flight": depth}

    # ---- roofline of the dominant kernel (events inside the C ABI around that kernel only) -------
    roof = None
    if rank == 0:
        pk = peaks()
        L.mb_profile_enable(1)
        ks = []
        for it in range(6):
            model.step(obs0, acts[it % H], indices=idx[it % H], noise=noise[it % H])
            ms = _lib.ctypes.c_float(0)
            _lib.check(L.mb_profile_last_ms(_lib.ctypes.byref(ms)))
            if it >= 2:
                ks.append(ms.value)
        L.mb_profile_enable(0)
        kms = float(np.mean(ks))
        achieved = FLOP_PER_TRANSITION * B / (kms * 1e-3) / 1e12
        roof = {"kernel": "k_pe_rollout_%s" % kernel, "bound": "tensor", "achieved": achieved,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"],
                "peak_source": pk["src"], "kernel_ms": kms, "launch_rows": B,
                "hbm_gbs_algorithmic": BYTES_PER_TRANSITION * B / (kms * 1e-3) / 1e9,
                "traffic": None}
        prof = os.path.join(ROOT, "profiles", "rollout_traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                roof["traffic"] = json.load(f).get(roof["kernel"])

    # ---- secondary metrics of BASELINE.json: model-train samples/s, REDQ / TQC targets/s ----------
    secondary = secondary_metrics(dev, world, rank, timed)

    if rank == 0:
        base, _ = cpu_arm(3, 1)
        launches_per_step = H * 2          # k_bucket + k_pe_rollout_tc per horizon step
        line = {"metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(workload_config(world), kernel=kernel),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "roofline": roof, "cpu_baseline": base, "secondary": secondary}
        print(json.dumps(line))
    pools.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
