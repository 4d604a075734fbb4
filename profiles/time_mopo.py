"""Timing of the all-elite paths: MOPO penalised step and the all-member forward."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

def timeit(fn, n=10, warm=3):
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

m, _ = make_pe_model(cases.MBPO, 11, "hopper")
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
for B in (50_000, 1_000_000):
    g = torch.Generator(device="cuda").manual_seed(0)
    o = torch.randn(B, 17, device="cuda", generator=g)
    a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
    eps = torch.randn(B, 18, device="cuda", generator=g)
    idx = torch.from_numpy(np.random.RandomState(0).choice(m.elite_indices, B).astype(np.uint8)).cuda()
    ms = timeit(lambda: m.step(o, a, indices=idx, noise=eps, return_distribution=True, penalty_type="max_var"))
    print("MOPO step (5 elites/row)  B=%7d: %.3f ms -> %.1f M transitions/s" % (B, ms, B / ms / 1e3))
    ms = timeit(lambda: m.predict_distribution(o, a))
    print("all-member forward (E=7)  B=%7d: %.3f ms -> %.1f M rows/s" % (B, ms, B / ms / 1e3))
    ms = timeit(lambda: m.step(o, a, indices=idx, noise=eps))
    print("plain step                B=%7d: %.3f ms" % (B, ms))
