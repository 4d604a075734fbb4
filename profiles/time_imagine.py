"""Where the time of one horizon step of B200MBPOCollector.imagine goes (1e5 rows, configs/mbpo.json shape):
policy action, rollout step with in-place pool insert, live-row compaction, loop-exit test."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import mirl_b200
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, make_pe_model
from tests.util import T
dev = "cuda"
env = FakeEnv("cheetah")
mc, _ = make_pe_model(cases.MBPO, 11, "cheetah", device=dev)
mc.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
pol = mirl_b200.B200GaussianPolicy(env, hidden_layers=[256, 256], activation="relu")
pW, pb = cases.mlp_params(91, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(pW, pb).items()})
pol = pol.to(dev)
N = 100_000
o = torch.randn(N, cases.OBS, device=dev)
block = torch.empty(N, mc.row_stride(), device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


a, _ = pol.action(o)
print("policy.action           %.3f ms" % timed(lambda: pol.action(o)))
print("model.step_rows         %.3f ms" % timed(lambda: mc.step_rows(o, a, block, row_offset=0, ordered=True)))
no, _, d = mc.step_rows(o, a, block, row_offset=0, ordered=True)
print("live-row compaction     %.3f ms" % timed(lambda: no[torch.nonzero(d.reshape(-1) < 0.5).reshape(-1)]))
print("loop-exit test (2 item) %.3f ms" % timed(lambda: (float(len(d)), float(torch.sum(d)))))
from mirl_b200.collectors import B200MBPOCollector
real = mirl_b200.B200DevicePool(env, max_size=200_000, device=dev)
real.add_samples(cases.dyn_samples(8001, 150_000))
for depth in (5,):
    imag = mirl_b200.B200DevicePool(env, max_size=10, device=dev, row_stride=mc.row_stride())
    col = B200MBPOCollector(mc, pol, real, imag, depth_schedule=depth, number_sample=N, save_n=2, bathch_size=N)
    np.random.seed(5)
    col.imagine(0)
    print("imagine depth %d         %.3f ms per call = %.3f ms per horizon step" % (depth, timed(lambda: col.imagine(0), 5), timed(lambda: col.imagine(0), 5) / depth))
    print("get_start_states        %.3f ms" % timed(lambda: col.get_start_states(), 5))
