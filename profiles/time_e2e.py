"""Where the end-to-end step spends its time (host draw, H2D, kernels, D2H) -- wall clock, synced."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

B = 1_000_000
m, _ = make_pe_model(cases.MBPO, 11, "cheetah")
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
h_o = torch.randn(B, 17).pin_memory()
h_a = (torch.rand(B, 6) * 2 - 1).pin_memory()
h_out = torch.empty(B, 19).pin_memory()

def wall(fn, tag, n=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print("%-34s %.3f ms" % (tag, (time.perf_counter() - t) / n * 1e3), flush=True)

o = h_o.cuda(); a = h_a.cuda()
wall(lambda: np.random.choice(m.elite_indices, B), "host np.random.choice")
wall(lambda: m.draw_members(B), "device draw_members (+state sync)")
wall(lambda: torch.randn(B, 18, device="cuda"), "torch.randn noise")
wall(lambda: h_o.to("cuda", non_blocking=True), "H2D obs 68 MB")
wall(lambda: h_a.to("cuda", non_blocking=True), "H2D act 24 MB")
idx = m.draw_members(B); eps = torch.randn(B, 18, device="cuda")
wall(lambda: m.step(o, a, indices=idx, noise=eps), "step (given idx, noise)")
no, r, d, _ = m.step(o, a, indices=idx, noise=eps)
wall(lambda: h_out.copy_(torch.cat([no, r, d], 1), non_blocking=True), "cat + D2H 76 MB")
m.device_draw_min_rows = 1 << 40
wall(lambda: m.step(o, a), "step() host draw")
m.device_draw_min_rows = 32768
wall(lambda: m.step(o, a), "step() device draw")
