"""Helpers shared by the tests: fixture loading, numpy<->torch packing, tolerances."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerance of record (BASELINE.json north_star: "within 1e-3 relative for fp32-accumulated
# outputs").  Element-wise relative error is meaningless next to zeros, so every comparison is
#   |got - ref| <= RTOL*|ref| + RTOL*mean|ref|        ("Precision").
RTOL = 1e-3


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def to_torch(p, device="cpu"):
    """numpy packed-parameter dict (tests/golden/cases.py) -> torch tensors."""
    out = {}
    for k, v in p.items():
        if isinstance(v, list):
            out[k] = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in v]
        elif isinstance(v, np.ndarray):
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device)
        else:
            out[k] = v
    return out


def T(x, device="cpu"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)


def assert_close(got, ref, rtol=RTOL, what=""):
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, dtype=np.float64)
    ref = np.asarray(ref.detach().cpu() if torch.is_tensor(ref) else ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s shape %s vs %s" % (what, got.shape, ref.shape)
    scale = np.abs(ref).mean() if ref.size else 0.0
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + rtol * scale + 1e-30
    worst = (err / bound).max() if ref.size else 0.0
    assert worst <= 1.0, "%s: max err/bound = %.3g (max abs err %.3g, scale %.3g)" % (
        what, worst, err.max(), scale)
    return err.max() if ref.size else 0.0


def norm_rel(got, ref):
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, dtype=np.float64)
    ref = np.asarray(ref.detach().cpu() if torch.is_tensor(ref) else ref, dtype=np.float64)
    return np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)


def proj(t, seed):
    a = (t.detach().cpu() if torch.is_tensor(t) else torch.as_tensor(t)).double().numpy()
    r = np.random.RandomState(seed).standard_normal(a.shape)
    return np.array([np.sqrt((a * a).sum()), (a * r).sum()])
