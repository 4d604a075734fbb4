"""D2H copy rates: contiguous vs the 2-D strided [next_obs | reward | done] slice of packed rows."""
import torch

B, O, A = 1_000_000, 17, 6
W = 2 * O + A + 2
dev = torch.device("cuda:0")
rows = torch.randn(B, W, device=dev)
flat = torch.randn(B, O + 2, device=dev)
h = torch.empty(B, O + 2).pin_memory()
hin = torch.randn(B, A).pin_memory()
din = torch.empty(B, A, device=dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = t(lambda: h.copy_(flat, non_blocking=True))
print("contiguous D2H 76 MB: %.3f ms = %.1f GB/s" % (ms, 76e6 / ms / 1e6))
ms = t(lambda: h.copy_(rows[:, O + A:], non_blocking=True))
print("strided    D2H 76 MB: %.3f ms = %.1f GB/s" % (ms, 76e6 / ms / 1e6))
ms = t(lambda: din.copy_(hin, non_blocking=True))
print("contiguous H2D 24 MB: %.3f ms = %.1f GB/s" % (ms, 24e6 / ms / 1e6))
s2 = torch.cuda.Stream()


def both():
    h.copy_(flat, non_blocking=True)
    with torch.cuda.stream(s2):
        din.copy_(hin, non_blocking=True)


ms = t(both)
torch.cuda.synchronize()
print("D2H 76 MB + H2D 24 MB on two streams: %.3f ms" % ms)
