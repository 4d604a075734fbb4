"""Timing of the model-train step and the REDQ / TQC targets (CUDA events, device-resident inputs)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import mirl_b200
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, make_critic, make_pe_model
from tests.util import T


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


m, _ = make_pe_model(cases.MBPO, 11)
tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, tb_log_freq=-1, silent=True, use_cuda_graph="--graph" in sys.argv)
for B in (256, 1024, 4096):
    bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, 7, B).items()}
    ms = timeit(lambda: tr.train_step_async(bt))
    print("train step  E=7 B=%5d: %.3f ms  -> %.2f M samples/s" % (B, ms, 7 * B / ms / 1e3))
g = torch.cuda.CUDAGraph()
bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, 7, 256).items()}
q, _ = make_critic("redq", 71)
t, _ = make_critic("tqc", 81)
for B in (256, 4096, 65536):
    x = cases.critic_inputs(72, B)
    o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
    r, d, lp = T(x["rewards"], "cuda"), T(x["terminals"], "cuda"), T(x["log_prob"], "cuda")
    ms = timeit(lambda: q.q_target(o, a, r, d, lp, alpha=0.2, sample_number=2, batchwise_sample=True))
    print("REDQ target B=%6d: %.3f ms -> %.2f M targets/s" % (B, ms, B / ms / 1e3))
    ms = timeit(lambda: t.atoms_target(o, a, r, d, lp, alpha=0.2, percentage_drop=8))
    print("TQC  target B=%6d: %.3f ms -> %.2f M targets/s" % (B, ms, B / ms / 1e3))
