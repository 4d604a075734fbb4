"""Profile driver: the persistent fused train kernel (one launch of 8 steps, E=7, B=256, bootstrap batches gathered
on the device), single-step launches, the fused critic update, the policy action and the critic targets."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import mirl_b200
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, make_critic, make_pe_model
from tests.util import T
m, _ = make_pe_model(cases.MBPO, 11)
tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, tb_log_freq=-1, silent=True)
N = 20_000
data = {k: T(v, "cuda") for k, v in cases.train_batch(5, 1, N, shared=True).items()}
index = torch.from_numpy(np.random.RandomState(1).randint(N, size=(7, N))).cuda()
for _ in range(3):
    tr.train_steps_async(data, index, N, 256, 0, 8)
bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, 7, 256).items()}
for _ in range(3):
    tr.train_step_async(bt)
q, _ = make_critic("redq", 71)
qt, _ = make_critic("redq", 171)
t, _ = make_critic("tqc", 81)
x = cases.critic_inputs(72, 256)
o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
r, d, lp = T(x["rewards"], "cuda"), T(x["terminals"], "cuda"), T(x["log_prob"], "cuda")
ct = mirl_b200.B200CriticTrainer(q, qt)
pol = mirl_b200.B200GaussianPolicy(FakeEnv(), hidden_layers=[256, 256], activation="relu").to("cuda")
for _ in range(2):
    ct.update_async(o, a, T(cases.critic_q_target(x), "cuda"))
    q.q_target(o, a, r, d, lp, alpha=0.2, sample_number=2, batchwise_sample=True)
    t.atoms_target(o, a, r, d, lp, alpha=0.2, percentage_drop=8)
    pol.action(torch.randn(100_000, 17, device="cuda"), return_log_prob=True)
torch.cuda.synchronize()
print("ok")
