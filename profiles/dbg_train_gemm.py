"""Debug: gradients of the train step with the tcgen05 GEMMs vs the SIMT GEMMs (two processes)."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
sys.path.insert(0, ".")

def run(tag):
    from mirl_b200 import _lib
    from tests.golden import cases
    from tests.helpers_gpu import make_pe_model
    from tests.util import T
    L = _lib.lib()
    out = {}
    for name, cfg, seed, E, B in (("small", cases.SMALL, 51, 3, 64), ("mbpo", cases.MBPO, 11, 7, 256), ("mbpo100", cases.MBPO, 11, 7, 100)):
        m, _ = make_pe_model(cfg, seed)
        bt = {k: T(v, "cuda") for k, v in cases.train_batch(41, E, B).items()}
        cm = m._c_model(None)
        tcfg = m.train_cfg(False)
        ws = torch.empty(L.mb_pe_train_workspace_bytes(ctypes.byref(cm), B), dtype=torch.uint8, device="cuda")
        grads = torch.zeros_like(m.module.flat.data)
        el = torch.zeros(m.ensemble_size, device="cuda")
        terms = torch.zeros(3, device="cuda")
        _lib.check(L.mb_pe_loss_grads(ctypes.byref(cm), ctypes.byref(tcfg), _lib.ptr(bt["observations"]),
                                      _lib.ptr(bt["actions"]), _lib.ptr(bt["deltas"]), _lib.ptr(bt["rewards"]),
                                      _lib.ptr(bt["terminals"]), B, _lib.ptr(grads), _lib.ptr(el), _lib.ptr(terms),
                                      _lib.ptr(ws), ws.numel(), _lib.stream()))
        torch.cuda.synchronize()
        out[name + "_g"] = grads.cpu().numpy(); out[name + "_el"] = el.cpu().numpy(); out[name + "_t"] = terms.cpu().numpy()
        segs = []
        d = m.module
        out[name + "_sizes"] = np.array([d.input_size] + list(d.hidden_layers) + [d.output_size]) if hasattr(d, "hidden_layers") else np.zeros(1)
    np.savez("gpurun_out/dbg_%s.npz" % tag, **out)

if len(sys.argv) > 1:
    run(sys.argv[1])
else:
    for tag, env in (("tc", {}), ("simt", {"MB_TRAIN_GEMM": "simt"})):
        subprocess.run([sys.executable, __file__, tag], env=dict(os.environ, **env), check=True)
    a, b = np.load("gpurun_out/dbg_tc.npz"), np.load("gpurun_out/dbg_simt.npz")
    for k in a.files:
        if k.endswith("_sizes"):
            continue
        x, y = a[k], b[k]
        err = np.abs(x - y)
        print("%-12s max|err| %.3g  max|ref| %.3g  rel %.3g  n_bad %d / %d  first_bad %s" % (
            k, err.max(), np.abs(y).max(), err.max() / (np.abs(y).max() + 1e-30), int((err > 1e-3 * (np.abs(y) + np.abs(y).mean())).sum()), x.size,
            np.argwhere(err > 1e-3 * (np.abs(y) + np.abs(y).mean()))[:5].ravel().tolist()))
        if k.endswith("_t") or k.endswith("_el"):
            print("    tc  ", x, "\n    simt", y)
