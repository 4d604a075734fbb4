"""Kernel-only timing of the rollout kernel via the C-ABI profiling hooks (mb_profile_*)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mirl_b200 import _lib
from tests.golden import cases
from tests.helpers_gpu import make_pe_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kernel = sys.argv[2] if len(sys.argv) > 2 else "tcgen05"
m, _ = make_pe_model(cases.MBPO, 11, "cheetah", kernel=kernel)
m.remember_loss([0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.4])
g = torch.Generator(device="cuda").manual_seed(0)
o = torch.randn(B, 17, device="cuda", generator=g)
a = torch.rand(B, 6, device="cuda", generator=g) * 2 - 1
eps = torch.randn(B, 18, device="cuda", generator=g)
idx = torch.from_numpy(np.random.RandomState(0).choice(m.elite_indices, B).astype(np.uint8)).cuda()
L = _lib.lib()
L.mb_profile_enable(1)
ts = []
for it in range(8):
    m.step(o, a, indices=idx, noise=eps)
    ms = ctypes.c_float(0)
    _lib.check(L.mb_profile_last_ms(ctypes.byref(ms)))
    ts.append(ms.value)
print("B=%d kernel=%s ms=%.4f  (%.1f M transitions/s)  all=%s" % (B, kernel, np.median(ts[2:]), B / np.median(ts[2:]) / 1e3, [round(t, 3) for t in ts]))
