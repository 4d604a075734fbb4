"""Profile driver: a few fused train steps (E=7, B=256) without graph replay; and critic targets."""
import sys
import torch
sys.path.insert(0, ".")
import mirl_b200
from tests.golden import cases
from tests.helpers_gpu import FakeEnv, make_critic, make_pe_model
from tests.util import T
m, _ = make_pe_model(cases.MBPO, 11)
tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, tb_log_freq=-1, silent=True, use_cuda_graph=False)
bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, 7, 256).items()}
for _ in range(4):
    tr.train_step_async(bt)
q, _ = make_critic("redq", 71)
t, _ = make_critic("tqc", 81)
x = cases.critic_inputs(72, 256)
o, a = T(x["obs"], "cuda"), T(x["act"], "cuda")
r, d, lp = T(x["rewards"], "cuda"), T(x["terminals"], "cuda"), T(x["log_prob"], "cuda")
for _ in range(2):
    q.q_target(o, a, r, d, lp, alpha=0.2, sample_number=2, batchwise_sample=True)
    t.atoms_target(o, a, r, d, lp, alpha=0.2, percentage_drop=8)
torch.cuda.synchronize()
print("ok")
