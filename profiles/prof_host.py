"""Host cost of one step_rows call (model draws members and noise itself): wall clock per call in
a free-running loop, and a cProfile of the same loop."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

m, _ = make_pe_model(cases.MBPO, 11, "cheetah")
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
for B in (125_000, 250_000, 1_000_000):
    o = torch.randn(B, 17, device="cuda"); a = torch.rand(B, 6, device="cuda") * 2 - 1
    buf = torch.empty(2, B, 42, device="cuda")

    def loop(n):
        for k in range(n):
            m.step_rows(o, a, buf[k & 1])
        torch.cuda.synchronize()

    for spec in (0, 1, 2, 3):
        m.draw_lookahead = spec
        loop(10)
        t = time.perf_counter(); loop(100); dt = (time.perf_counter() - t) / 100
        print("B=%7d spec=%d: %.3f ms per call (%.1f M rows/s)" % (B, spec, dt * 1e3, B / dt / 1e6), flush=True)
    if B == 125_000:
        pr = cProfile.Profile(); pr.enable(); loop(100); pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(14)
