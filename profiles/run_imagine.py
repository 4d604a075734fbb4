import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import mirl_b200
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, make_pe_model
from tests.util import T
from mirl_b200.collectors import B200MBPOCollector
dev = "cuda"
env = FakeEnv("cheetah")
mc, _ = make_pe_model(cases.MBPO, 11, "cheetah", device=dev)
mc.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
pol = mirl_b200.B200GaussianPolicy(env, hidden_layers=[256, 256], activation="relu")
pW, pb = cases.mlp_params(91, 1, [cases.OBS, 256, 256, 2 * cases.ACT])
pol.load_state_dict({k: T(v) for k, v in cases.policy_state_dict(pW, pb).items()})
pol = pol.to(dev)
N = 100_000
real = mirl_b200.B200DevicePool(env, max_size=200_000, device=dev)
real.add_samples(cases.dyn_samples(8001, 150_000))
imag = mirl_b200.B200DevicePool(env, max_size=10, device=dev, row_stride=mc.row_stride())
col = B200MBPOCollector(mc, pol, real, imag, depth_schedule=2, number_sample=N, save_n=2, bathch_size=N)
np.random.seed(5)
col.imagine(0)
col.imagine(0)
torch.cuda.synchronize()
print("ok")
