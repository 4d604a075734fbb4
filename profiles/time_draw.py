"""Kernel-only time of the device member draw (CUDA events) and the host-side pieces around it."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mirl_b200 import _lib
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
st = np.random.get_state()
host = np.zeros(627, dtype=np.uint32); host[:624], host[624] = st[1], st[2]
state = torch.from_numpy(host.view(np.int32)).cuda()
out = torch.empty(B, dtype=torch.uint8, device="cuda")
ws = torch.empty(L.mb_pe_draw_members_workspace_bytes(B, 5), dtype=torch.uint8, device="cuda")
for n_el in (5, 7):
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.mb_pe_draw_members(_lib.ptr(state), bytes(range(n_el)), n_el, B, _lib.ptr(out),
                                        _lib.ptr(ws), ws.numel(), _lib.stream()))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("draw kernel n_elites=%d B=%d: %s ms" % (n_el, B, [round(t, 3) for t in ts]))
t = time.perf_counter()
for _ in range(20):
    s = np.random.get_state(); np.random.set_state(s)
print("get_state+set_state %.3f ms" % ((time.perf_counter() - t) / 20 * 1e3))
t = time.perf_counter()
for _ in range(20):
    x = torch.from_numpy(host.view(np.int32)).cuda(); y = x.cpu()
print("H2D+D2H 2.5KB %.3f ms" % ((time.perf_counter() - t) / 20 * 1e3))
