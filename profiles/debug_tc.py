"""Diagnostic (not a test): tcgen05 vs SIMT rollout kernels on one small batch."""
import sys
import numpy as np
import torch
from tests.golden import cases
from tests.helpers_gpu import make_pe_model
from tests.util import T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ms, _ = make_pe_model(cases.MBPO, 11, "cheetah", kernel="simt")
mt, _ = make_pe_model(cases.MBPO, 11, "cheetah", kernel="tcgen05")
for m in (ms, mt):
    m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
o, a = cases.rollout_inputs(5, B)
rs = np.random.RandomState(1)
idx = rs.choice(ms.elite_indices, B)
eps = rs.standard_normal((B, 18)).astype(np.float32)
o, a, eps = T(o, "cuda"), T(a, "cuda"), T(eps, "cuda")
n1, r1, d1, _ = ms.step(o, a, indices=idx, noise=eps)
torch.cuda.synchronize()
print("simt done", flush=True)
n2, r2, d2, _ = mt.step(o, a, indices=idx, noise=eps)
torch.cuda.synchronize()
err = (n1 - n2).abs()
print("B=%d max|err| next_obs %.3e  reward %.3e  scale %.3e" % (B, err.max().item(), (r1 - r2).abs().max().item(), n1.abs().mean().item()))
print("rows with err>1e-3:", int((err.max(1)[0] > 1e-3).sum().item()), "of", B)
print("per-col max err:", [round(v, 5) for v in err.max(0)[0].tolist()])
print("first rows err:", [round(v, 5) for v in err.max(1)[0][:16].tolist()])
