"""Kernel-only timing of step() vs step_rows() (contiguous / chained strided obs) via mb_profile_*."""
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
rows = torch.zeros(2 * B, 42, device="cuda")
L = _lib.lib()
L.mb_profile_enable(1)

def t(fn, tag):
    ts = []
    for it in range(8):
        fn()
        ms = ctypes.c_float(0)
        _lib.check(L.mb_profile_last_ms(ctypes.byref(ms)))
        ts.append(ms.value)
    print("%-28s ms=%.4f" % (tag, np.median(ts[2:])), flush=True)

t(lambda: m.step(o, a, indices=idx, noise=eps), "step")
t(lambda: m.step_rows(o, a, rows, indices=idx, noise=eps), "step_rows contiguous obs")
v = rows[:B, 23:40]
t(lambda: m.step_rows(v, a, rows, row_offset=B, indices=idx, noise=eps), "step_rows strided obs")
