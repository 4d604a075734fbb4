"""Does the member draw (side stream) overlap a rollout in flight?  Wall clock of draw_members()
issued right behind an un-synchronised rollout launch, vs alone; with 0/1/2 SMs kept free."""
import sys, time
import torch
sys.path.insert(0, ".")
from mirl_b200 import _lib
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

B = 1_000_000
m, _ = make_pe_model(cases.MBPO, 11, "cheetah")
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
o = torch.randn(B, 17, device="cuda"); a = torch.rand(B, 6, device="cuda") * 2 - 1
idx = m.draw_members(B); eps = torch.randn(B, 18, device="cuda")
for n in (B, B // 4):
    for res in (0, 1, 2):
        _lib.lib().mb_set_reserved_sms(res)
        for _ in range(3):
            m.step(o, a, indices=idx, noise=eps); m.draw_members(n)
        torch.cuda.synchronize()
        alone = behind = total = 0.0
        for _ in range(10):
            t = time.perf_counter(); m.draw_members(n); alone += time.perf_counter() - t
            torch.cuda.synchronize()
            t = time.perf_counter()
            m.step(o, a, indices=idx, noise=eps)
            t1 = time.perf_counter(); m.draw_members(n); behind += time.perf_counter() - t1
            torch.cuda.synchronize(); total += time.perf_counter() - t
        print("draw %7d rows, reserved SMs %d: alone %.3f ms, behind a rollout %.3f ms, rollout+draw %.3f ms"
              % (n, res, alone * 100, behind * 100, total * 100), flush=True)
