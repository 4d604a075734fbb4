"""Profile driver: a few tcgen05 rollout steps at B rows (used under ncu; see profiles/README.md)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kernel = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
m, _ = make_pe_model(cases.MBPO, 11, "cheetah", kernel=kernel)
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
g = torch.Generator(device="cuda").manual_seed(0)
o = torch.randn(B, 17, device="cuda", generator=g)
a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
eps = torch.randn(B, 18, device="cuda", generator=g)
idx = torch.from_numpy(np.random.RandomState(0).choice(m.elite_indices, B).astype(np.uint8)).cuda()
for _ in range(steps):
    no, r, d, _ = m.step(o, a, indices=idx, noise=eps)
torch.cuda.synchronize()
print("ok", float(no.abs().mean()))
