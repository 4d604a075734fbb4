"""Debug: fused training forward vs per-layer forward, each against a float64 evaluation of the oracle."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mirl_b200 import _lib
from oracle import pe_oracle as po
from tests.golden import cases
from tests.helpers_gpu import make_pe_model
from tests.util import T, to_torch

obs, act, E, B = int(sys.argv[1]), int(sys.argv[2]), 7, 130
cfg = dict(E=E, elites=5, hidden=[200, 200, 200, 200], act="swish")
m, p = make_pe_model(cfg, 61, "cheetah", kernel="tcgen05", obs=obs, act=act)
bt = cases.train_batch(64, E, B, obs=obs, act=act)
pt = to_torch(p)
pd = {k: ([w.double() for w in v] if isinstance(v, list) else (v.double() if torch.is_tensor(v) else v)) for k, v in pt.items()}
ref = po.model_loss_grads(pd, *[T(bt[k]).double() for k in ("observations", "actions", "deltas", "rewards", "terminals")],
                          ignore_terminal_state=False, weight_decay=cases.WEIGHT_DECAY)
L = _lib.lib()
btc = {k: T(v, "cuda") for k, v in bt.items()}
for tag, pack in (("per-layer", None), ("fused", torch.empty(L.mb_pe_tc_pack_bytes(ctypes.byref(m._c_model(None))), dtype=torch.uint8, device="cuda"))):
    cm = m._c_model(pack)
    tcfg = m.train_cfg(False)
    ws = torch.empty(L.mb_pe_train_workspace_bytes(ctypes.byref(cm), B), dtype=torch.uint8, device="cuda")
    grads = torch.zeros_like(m.module.flat.data)
    el = torch.zeros(E, device="cuda"); terms = torch.zeros(3, device="cuda")
    _lib.check(L.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(tcfg), _lib.ptr(btc["observations"]), _lib.ptr(btc["actions"]),
                                  _lib.ptr(btc["deltas"]), _lib.ptr(btc["rewards"]), _lib.ptr(btc["terminals"]), B,
                                  _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms), _lib.ptr(ws), ws.numel(), _lib.stream()))
    torch.cuda.synchronize()
    errs = []
    for l in range(5):
        g = m.module.weight(l, flat=grads).cpu().double()
        r = ref["W"][l]
        errs.append(float((g - r).norm() / r.norm()))
    print("%-10s loss rel err %.3g   dW rel err per layer %s" % (tag, abs(float(terms.sum()) - float(ref["loss"])) / abs(float(ref["loss"])), ["%.2g" % e for e in errs]))

# ---- per-layer pre-activations saved in the workspace: [x0 | z_0 .. z_{L-2} | ...]
def ru4(n):
    return (n + 3) // 4 * 4
IN = obs + act
x = torch.cat([po.normalize(T(bt["observations"]).double(), pd["obs_mean"], pd["obs_std"]),
               po.normalize(T(bt["actions"]).double(), pd["act_mean"], pd["act_std"])], -1)
zs_ref = []
h = x
for l in range(5):
    z = h.matmul(pd["W"][l]) + pd["b"][l]
    zs_ref.append(z)
    h = z * torch.sigmoid(z)
for tag, pack in (("per-layer", None), ("fused", torch.empty(L.mb_pe_tc_pack_bytes(ctypes.byref(m._c_model(None))), dtype=torch.uint8, device="cuda"))):
    cm = m._c_model(pack)
    tcfg = m.train_cfg(False)
    ws = torch.zeros(L.mb_pe_train_workspace_bytes(ctypes.byref(cm), B), dtype=torch.uint8, device="cuda")
    grads = torch.zeros_like(m.module.flat.data)
    el = torch.zeros(E, device="cuda"); terms = torch.zeros(3, device="cuda")
    _lib.check(L.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(tcfg), _lib.ptr(btc["observations"]), _lib.ptr(btc["actions"]),
                                  _lib.ptr(btc["deltas"]), _lib.ptr(btc["rewards"]), _lib.ptr(btc["terminals"]), B,
                                  _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms), _lib.ptr(ws), ws.numel(), _lib.stream()))
    torch.cuda.synchronize()
    f = ws.view(torch.float32)
    off = ru4(E * B * IN)
    out = []
    for l in range(4):
        n = E * B * 200
        z = f[off:off + n].view(E, B, 200).cpu().double()
        off += ru4(n)
        out.append("%.2g" % float((z - zs_ref[l]).abs().max() / zs_ref[l].abs().max()))
    print("%-10s max|dz|/max|z| per hidden layer: %s" % (tag, out))
    off += ru4(E * B * 200)                      # skip gA
    D2 = 2 * (obs + 1)
    head = f[off:off + E * B * D2].view(E, B, D2).cpu().double()
    ref_h = zs_ref[4]
    d = (head - ref_h).abs()
    print("%-10s head: max|d mean| %.3g (max|mean| %.3g)   max|d rawls| %.3g (max|rawls| %.3g)" % (
        tag, float(d[..., :obs + 1].max()), float(ref_h[..., :obs + 1].abs().max()),
        float(d[..., obs + 1:].max()), float(ref_h[..., obs + 1:].abs().max())))
