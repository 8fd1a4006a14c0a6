"""Sweep tile width x {single, pair} over the network's real GEMM/conv shapes (8-view and 1-view batches) and print the
best choice next to what the built-in heuristic (bn=0) picks.  Output feeds choose_tile() in tc_gemm.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreammat_b200 import dense_ops as D
from dreammat_b200._cabi import lib
dev = "cuda"
def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
lib().dm_tune_gemm(11)
shapes = []
for B in (24, 3):
    for hw, c in ((64, 320), (32, 640), (16, 1280), (8, 1280)):
        shapes.append(("conv", B, hw, c, c))
        if c != 320 or True:
            shapes.append(("conv", B, hw, 2 * c if hw != 64 else 640, c))
        M = B * hw * hw
        if hw != 8:
            shapes += [("gemm", M, c, c), ("gemm", M, 3 * c, c), ("gemm", M, 8 * c, c), ("gemm", M, c, 4 * c)]
    shapes.append(("conv", B, 32, 960, 640)); shapes.append(("conv", B, 16, 1920, 1280))
for B in (8, 1):
    shapes += [("conv", B, 512, 128, 128), ("conv", B, 256, 128, 256), ("conv", B, 256, 256, 256), ("conv", B, 128, 256, 512),
               ("conv", B, 128, 512, 512), ("conv", B, 64, 512, 512)]
for sh in shapes:
    if sh[0] == "conv":
        _, n, hw, ci, co = sh
        x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
        fn = lambda bn: D.conv2d(x, w, 3, bn=bn); fl = 2 * n * hw * hw * ci * co * 9; N = co
        name = f"conv {n}x{hw}^2 {ci}->{co}"
    else:
        _, m, nn, k = sh
        a = torch.randn(m, k, device=dev).half(); b = (torch.randn(nn, k, device=dev) * 0.05).half()
        fn = lambda bn: D.gemm(a, b, bn=bn); fl = 2 * m * nn * k; N = nn
        name = f"gemm {m}x{nn}x{k}"
    cands = [64, 128] + ([256] if N >= 256 else []) + [1128] + ([1160] if N % 160 == 0 else []) + ([1256] if N >= 256 else [])
    res = {}
    for bn in [0] + cands:
        try:
            res[bn] = timeit(lambda: fn(bn))
        except Exception as ex:  # noqa: BLE001
            res[bn] = float("inf")
    best = min(cands, key=lambda b: res[b])
    tag = lambda b: ("p" if b >= 1000 else "s") + str(b % 1000)
    print(f"{name:30s} auto {res[0]*1e3:8.1f} us | best {tag(best):5s} {res[best]*1e3:8.1f} us {fl/res[best]/1e9:7.1f} TF/s | "
          + " ".join(f"{tag(b)}:{res[b]*1e3:.1f}" for b in cands) + ("   <-- auto off by %.0f%%" % (100 * (res[0] / res[best] - 1)) if res[0] > 1.03 * res[best] else ""), flush=True)
