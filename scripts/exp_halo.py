"""Halo-reuse 3x3 convolution experiment (dm_tune_gemm 31 / 32): correctness of both descriptor modes vs fp32 torch, then
timing against the shipped pair / single-CTA kernels on the VAE's wide layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dreammat_b200 import dense_ops as D
from dreammat_b200._cabi import lib
dev = "cuda"
def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
g = torch.Generator(device=dev).manual_seed(0)
good = {}
for mode in (31, 32):
    lib().dm_tune_gemm(mode)
    errs = []
    for (n, hw, ci, co) in ((2, 128, 64, 128), (1, 256, 128, 256), (2, 128, 128, 512)):
        x = torch.randn(n, hw, hw, ci, device=dev, generator=g).half()
        w = (torch.randn(co, ci, 3, 3, device=dev, generator=g) / 30).half()
        bias = torch.randn(co, device=dev, generator=g).half()
        y = D.conv2d(x, D.conv_weight_to_gemm(w), 3, bias=bias)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1) + bias.float()[None, :, None, None]
        errs.append(rel(y.float().permute(0, 3, 1, 2), ref))
    good[mode] = max(errs) < 1e-3
    print(f"mode {mode}: rel errs {['%.2e' % e for e in errs]} -> {'OK' if good[mode] else 'WRONG'}", flush=True)
print("== timing", flush=True)
for (n, hw, ci, co) in ((8, 512, 128, 128), (8, 256, 128, 256), (8, 256, 256, 256), (8, 128, 256, 512), (8, 128, 512, 512)):
    x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
    fl = 2 * n * hw * hw * ci * co * 9
    r = []
    for mode in (30, 31, 32):
        if mode != 30 and not good[mode]:
            continue
        lib().dm_tune_gemm(mode)
        ms = timeit(lambda: D.conv2d(x, w, 3))
        r.append(f"mode{mode}: {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s")
    print(f"conv {n}x{hw}^2 {ci}->{co}: " + " | ".join(r), flush=True)
lib().dm_tune_gemm(30)
