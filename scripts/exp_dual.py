"""Two accumulators per tile in the single-CTA GEMM kernel (dm_tune_gemm 41): correctness vs fp32 torch, then timing
against one accumulator on 64- and 128-wide tiles (tests whether the ~150-cycle-per-MMA floor is the accumulator chain)."""
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
lib().dm_tune_gemm(10); lib().dm_tune_gemm(20); lib().dm_tune_gemm(41)      # no pairs, no split-K, dual on
for (M, N, K, bn) in ((300, 200, 192, 64), (1000, 320, 320, 128), (4096, 640, 1024, 128), (20000, 128, 576, 64)):
    a = torch.randn(M, K, device=dev, generator=g).half(); b = (torch.randn(N, K, device=dev, generator=g) * 0.05).half()
    bias = torch.randn(N, device=dev, generator=g).half(); res = torch.randn(M, N, device=dev, generator=g).half()
    out = D.gemm(a, b, bias=bias, residual=res, act="silu", bn=bn)
    ref = F.silu(a.float() @ b.float().t() + bias.float()) + res.float()
    print(f"dual gemm {M}x{N}x{K} bn={bn}: rel {rel(out, ref):.2e}", flush=True)
x = torch.randn(2, 64, 64, 128, device=dev, generator=g).half(); w = (torch.randn(128, 128, 3, 3, device=dev, generator=g) / 34).half()
y = D.conv2d(x, D.conv_weight_to_gemm(w), 3, bn=128)
print(f"dual conv: rel {rel(y.float().permute(0, 3, 1, 2), F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1)):.2e}", flush=True)
cases = []
def conv(n, hw, ci, co, bn):
    x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
    cases.append((f"conv {n}x{hw}^2 {ci}->{co} bn={bn}", lambda: D.conv2d(x, w, 3, bn=bn), 2 * n * hw * hw * ci * co * 9))
def gemm(m, n, k, bn):
    a = torch.randn(m, k, device=dev).half(); b = (torch.randn(n, k, device=dev) * 0.05).half()
    cases.append((f"gemm {m}x{n}x{k} bn={bn}", lambda: D.gemm(a, b, bn=bn), 2 * m * n * k))
conv(8, 512, 128, 128, 128); conv(8, 512, 64, 64, 64); gemm(8192, 4096, 4096, 128); gemm(8192, 4096, 4096, 64); conv(24, 8, 2560, 1280, 64)
conv(24, 64, 320, 320, 128)
for name, fn, fl in cases:
    r = []
    for code in (40, 41):
        lib().dm_tune_gemm(code)
        ms = timeit(fn)
        r.append(f"{'dual' if code == 41 else 'one '}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s")
    print(f"{name:34s} " + " | ".join(r), flush=True)
lib().dm_tune_gemm(40); lib().dm_tune_gemm(11); lib().dm_tune_gemm(21)
