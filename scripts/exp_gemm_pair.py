"""Correctness + A/B of the CTA-pair (cta_group::2) GEMM / conv kernel against the single-CTA kernel.
First runs small checks under a watchdog-friendly order (a hang shows before the long list starts)."""
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
print("== correctness (pair forced)", flush=True)
lib().dm_tune_gemm(12)
for (M, N, K, bn) in ((256, 128, 64, 128), (256, 256, 128, 256), (300, 200, 192, 128), (1000, 320, 320, 128), (4096, 1280, 640, 256),
                      (98304, 320, 320, 128), (513, 768, 1024, 256)):
    a = torch.randn(M, K, device=dev, generator=g).half(); b = (torch.randn(N, K, device=dev, generator=g) * 0.05).half()
    bias = torch.randn(N, device=dev, generator=g).half(); res = torch.randn(M, N, device=dev, generator=g).half()
    out = D.gemm(a, b, bias=bias, residual=res, act=None, bn=bn)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias.float() + res.float()
    print(f"gemm {M}x{N}x{K} bn={bn}: rel {rel(out, ref):.2e}", flush=True)
for (n, hw, ci, co, bn, stride) in ((2, 32, 128, 192, 128, 1), (1, 16, 64, 256, 256, 1), (3, 8, 128, 128, 128, 1), (2, 64, 64, 128, 128, 2), (8, 64, 320 + 0, 320, 128, 1)):
    ci_p = (ci + 63) // 64 * 64
    x = torch.randn(n, hw, hw, ci_p, device=dev, generator=g).half()
    w = (torch.randn(co, ci_p, 3, 3, device=dev, generator=g) / 30).half()
    if stride == 1:
        y = D.conv2d(x, D.conv_weight_to_gemm(w), 3, bn=bn)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1)
    else:
        y = D.conv2d(x, D.conv_weight_to_gemm(w), 3, stride=2, pad=(0, 0), out_hw=(hw // 2, hw // 2), bn=bn)
        ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), stride=2)
    torch.cuda.synchronize()
    print(f"conv {n}x{hw}^2 {ci}->{co} bn={bn} s{stride}: rel {rel(y.float().permute(0, 3, 1, 2), ref):.2e}", flush=True)
# geglu epilogue through the pair kernel
a = torch.randn(4096, 320, device=dev, generator=g).half(); w = (torch.randn(2560, 320, device=dev, generator=g) / 18).half()
bias = torch.randn(2560, device=dev, generator=g).half()
out = D.gemm(a, D.geglu_interleave(w), bias=D.geglu_interleave(bias), act="geglu", bn=256)
pr = a.float() @ w.float().t() + bias.float()
print(f"geglu pair: rel {rel(out, pr[:, :1280] * F.gelu(pr[:, 1280:])):.2e}", flush=True)

print("== timing: single-CTA vs pair", flush=True)
cases = []
def conv(n, hw, ci, co, bn=0):
    x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
    cases.append((f"conv {n}x{hw}^2 {ci}->{co} bn={bn}", lambda: D.conv2d(x, w, 3, bn=bn), 2 * n * hw * hw * ci * co * 9))
def gemm(m, n, k, bn=0):
    a = torch.randn(m, k, device=dev).half(); b = (torch.randn(n, k, device=dev) * 0.05).half()
    cases.append((f"gemm {m}x{n}x{k} bn={bn}", lambda: D.gemm(a, b, bn=bn), 2 * m * n * k))
conv(8, 512, 128, 128); conv(8, 256, 256, 256, 128); conv(8, 256, 256, 256, 256); conv(8, 128, 512, 512, 128); conv(8, 128, 512, 512, 256)
conv(24, 64, 320, 320, 128); conv(24, 64, 640, 320, 128); conv(24, 32, 640, 640, 128); conv(24, 32, 640, 640, 256); conv(24, 32, 1280, 640, 128)
conv(24, 16, 1280, 1280, 128); conv(24, 16, 1280, 1280, 256); conv(24, 16, 2560, 1280, 128); conv(24, 8, 2560, 1280, 128)
conv(3, 64, 320, 320, 128); conv(3, 32, 640, 640, 128); conv(3, 16, 1280, 1280, 128)
gemm(8192, 4096, 4096, 256); gemm(8192, 4096, 4096, 128)
gemm(98304, 320, 320, 128); gemm(98304, 960, 320, 128); gemm(98304, 2560, 320, 256); gemm(98304, 2560, 320, 128); gemm(98304, 320, 1280, 128)
gemm(24576, 5120, 640, 256); gemm(24576, 640, 2560, 128); gemm(6144, 10240, 1280, 256); gemm(6144, 1280, 5120, 128); gemm(6144, 1280, 5120, 256)
for name, fn, fl in cases:
    r = []
    for code, tag in ((10, "single"), (12, "pair")):
        lib().dm_tune_gemm(code)
        ms = timeit(fn)
        r.append(f"{tag}: {ms:.3f} ms {fl/ms/1e9:7.1f} TF/s")
    print(f"{name:38s} " + " | ".join(r), flush=True)
