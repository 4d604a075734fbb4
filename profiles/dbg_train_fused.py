"""Debug / timing driver for the persistent fused training kernel (train_fused.cu).
Run on the GPU box:  python profiles/dbg_train_fused.py [check] [time]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mirl_b200  # noqa: E402
from oracle import pe_oracle as po  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.helpers_gpu import FakeEnv, make_pe_model  # noqa: E402
from tests.util import T, norm_rel, to_torch  # noqa: E402


def check(cfg, seed, B):
    m, p = make_pe_model(cfg, seed)
    E = cfg["E"]
    tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, lr=1e-3, tb_log_freq=-1, silent=True)
    bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, E, B).items()}
    el, terms = tr.train_step_async(bt)
    torch.cuda.synchronize()
    pt = to_torch(p)
    g = po.model_loss_grads(pt, *[bt[k].cpu() for k in ("observations", "actions", "deltas", "rewards", "terminals")],
                            weight_decay=cases.WEIGHT_DECAY[:len(cfg["hidden"]) + 1])
    print("cfg E=%d hidden=%s B=%d" % (E, cfg["hidden"], B))
    print("  ensemble_loss got", el.cpu().numpy()[:4], "\n                ref", g["ensemble_loss"].numpy()[:4])
    print("  terms", terms.cpu().numpy(), "ref loss", float(g["loss"]))
    msd = m.module.reference_state_dict("module.", flat=tr.exp_avg)
    L = len(cfg["hidden"]) + 1
    for l in range(L):
        errs = [norm_rel(msd["module.net.%d.weight_net%d" % (2 * l, e)].cpu().numpy() / 0.1, g["W"][l][e].numpy())
                for e in range(E)]
        errb = [norm_rel(msd["module.net.%d.bias_net%d" % (2 * l, e)].cpu().numpy() / 0.1, g["b"][l][e].numpy())
                for e in range(E)]
        print("  layer %d  dW rel err max %.3g   db rel err max %.3g" % (l, max(errs), max(errb)))
    print("  dmaxls rel err", max(norm_rel(msd["module.maxlogstd_net%d" % e].cpu().numpy() / 0.1,
                                            g["max_logstd"][e].numpy()) for e in range(E)),
          " dminls", max(norm_rel(msd["module.minlogstd_net%d" % e].cpu().numpy() / 0.1,
                                  g["min_logstd"][e].numpy()) for e in range(E)))


def timing(B=256, nsteps=400, members=None):
    cfg = cases.MBPO
    m, _ = make_pe_model(cfg, 11)
    kw = {} if members is None else {"members": members}
    tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, lr=1e-3, tb_log_freq=-1, silent=True, **kw)
    N = 100_000
    data = {k: T(v, "cuda") for k, v in cases.train_batch(5, 1, N, shared=True).items()}
    index = torch.from_numpy(np.random.RandomState(1).randint(N, size=(7, N))).cuda()
    tr.train_steps_async(data, index, N, B, 0, 20)
    torch.cuda.synchronize()
    ne = (members[1] - members[0]) if members else 7
    for n in (1, 40, nsteps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if n == 1 else 3
        ev0.record()
        for _ in range(reps):
            tr.train_steps_async(data, index, N, B, 0, n)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        print("  members %s B=%d: launch of %4d steps: %.3f ms  = %.2f us/step  (%.1f M samples/s)" %
              (members or (0, 7), B, n, ms, 1e3 * ms / n, ne * B * n / ms / 1e3))
    bt = {k: T(v, "cuda") for k, v in cases.train_batch(600, 7, B).items()}
    for _ in range(5):
        tr.train_step_async(bt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        tr.train_step_async(bt)
    torch.cuda.synchronize()
    print("  train_step_async (one launch per step, host loop): %.2f us/step" % ((time.perf_counter() - t0) / 200 * 1e6))


if __name__ == "__main__":
    what = sys.argv[1:] or ["check", "time"]
    print("MB_TRAIN_FUSED =", os.environ.get("MB_TRAIN_FUSED"))
    if "check" in what:
        check(cases.SMALL, 51, 64)
        check(cases.MBPO, 11, 256)
        check(cases.MBPO, 11, 37)
        check(cases.MBPO, 11, 1)
        check(cases.MBPO, 11, 600)
    if "time" in what:
        timing()
        timing(members=(0, 1))
        timing(B=1024, nsteps=100)


def timeline():
    """Needs the -DMB_FUSED_TIMELINE build (profiles/build_fused_variant.sh tl -DMB_FUSED_TIMELINE, selected
    with MIRL_B200_LIB): CTA 0's clock64 stamps of step 2."""
    cfg = cases.MBPO
    m, _ = make_pe_model(cfg, 11)
    tr = mirl_b200.B200PEModelTrainer(FakeEnv(), m, lr=1e-3, tb_log_freq=-1, silent=True)
    N, B = 100_000, 256
    data = {k: T(v, "cuda") for k, v in cases.train_batch(5, 1, N, shared=True).items()}
    index = torch.from_numpy(np.random.RandomState(1).randint(N, size=(7, N))).cuda()
    tr.train_steps_async(data, index, N, B, 0, 4)
    torch.cuda.synchronize()
    tr.train_steps_async(data, index, N, B, 0, 4)
    torch.cuda.synchronize()
    d = tr._ws[4096:8192].view(torch.int64).cpu().numpy()
    t0 = d[0]
    print("timeline of CTA 0, step 2 (cycles from step start):")
    print("  phase A end %d | barrier1 passed %d | phase B end %d | barrier2 passed %d" % tuple(d[1:5] - t0))
    print("  pass starts:", list(d[8:17] - t0))
    print("  job epilogue starts %s" % (list(d[40:44] - t0)))
    g0 = d[200:221].min()
    print("  wall clock (ns from the first CTA's step start), member 0, per slot:")
    for name, o in (("phase A start", 200), ("phase A end  ", 230), ("phase B start", 290), ("phase B end  ", 260)):
        print("    %s" % name, " ".join("%6d" % v for v in (d[o:o + 21] - g0)))
    print("  per pass: start | after global copy | after pad zero | first chunk landed | chunks done | epilogue done | synced")
    for p in range(9):
        print("   pass %d:" % p, d[8 + p] - t0, list(d[60 + 6 * p:66 + 6 * p] - t0))


if __name__ == "__main__" and "timeline" in sys.argv[1:]:
    timeline()
