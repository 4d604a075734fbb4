import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import mirl_b200
from oracle import critic_oracle as co
from tests.golden import cases
from tests.helpers_gpu import FakeEnv
from tests.util import T
pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu")
pW, pb = cases.mlp_params(91, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(pW, pb).items()})
pol = pol.to("cuda")
for B in (4096, 5000, 100_000):
    o, _ = cases.rollout_inputs(7, B)
    eps = np.random.RandomState(9).standard_normal((B, cases.ACT)).astype(np.float32)
    act, info = pol.action(T(o, "cuda"), return_log_prob=True, return_mean_std=True, noise=T(eps, "cuda"))
    torch.cuda.synchronize()
    n = min(B, 20000)
    oa, olp, om, os_ = co.policy_action([T(a) for a in pW], [T(a) for a in pb], T(o[:n]), T(eps[:n]))
    print(B, "max|err| action %.2e  mean %.2e  std rel %.2e  log_prob %.2e" % (
        float((act[:n].cpu() - oa).abs().max()), float((info["mean"][:n].cpu() - om).abs().max()),
        float(((info["std"][:n].cpu() - os_) / os_).abs().max()), float((info["log_prob"][:n].cpu() - olp).abs().max())))
o = torch.randn(100_000, cases.OBS, device="cuda")
for _ in range(3): pol.action(o)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): pol.action(o)
torch.cuda.synchronize(); print("policy.action 1e5 rows: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
noise = torch.randn(100_000, cases.ACT, device="cuda")
for _ in range(3): pol.action(o, noise=noise)
torch.cuda.synchronize()
ev0.record()
for _ in range(50): pol.action(o, noise=noise)
ev1.record()
torch.cuda.synchronize()
print("policy.action 1e5 rows, given noise, CUDA events: %.3f ms per call (GPU busy time incl. gaps)" % (ev0.elapsed_time(ev1) / 50))
