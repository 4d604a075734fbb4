import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import pe_oracle as po
from tests.golden import cases
from tests.helpers_gpu import make_pe_model
from tests.util import T, to_torch
B = 20000
m, p = make_pe_model(cases.MBPO, 11, "cheetah", kernel="tcgen05")
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
o, a = cases.rollout_inputs(5, B)
eps = np.random.RandomState(6).standard_normal((B, 18)).astype(np.float32)
idx = np.random.RandomState(7).choice(m.elite_indices, B)
no, r, d, _ = m.step(T(o, "cuda"), T(a, "cuda"), indices=idx, noise=T(eps, "cuda"))
p64 = {k: ([t.double() for t in v] if isinstance(v, list) else (v.double() if torch.is_tensor(v) else v)) for k, v in to_torch(p).items()}
rno, rr, rd = po.rollout_step(p64, T(o).double(), T(a).double(), torch.from_numpy(idx), T(eps).double(), "cheetah")
e = (no.cpu().double() - rno).abs()
print("next_obs: max abs err %.3g  mean|ref| %.3g  worst err/(1e-3|ref| + 1e-3 mean) %.3g" % (e.max(), rno.abs().mean(), (e / (1e-3 * rno.abs() + 1e-3 * rno.abs().mean())).max()))
e = (r.cpu().double() - rr).abs()
print("reward:   max abs err %.3g  mean|ref| %.3g  worst ratio %.3g" % (e.max(), rr.abs().mean(), (e / (1e-3 * rr.abs() + 1e-3 * rr.abs().mean())).max()))
