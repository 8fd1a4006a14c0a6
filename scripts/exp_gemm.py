"""A/B the persistent GEMM's CTAs-per-SM setting on the shapes the networks actually use."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreammat_b200 import dense_ops as D
from dreammat_b200._cabi import lib
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
cases = []
def conv(n, hw, ci, co, bn=0):
    x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
    cases.append((f"conv {n}x{hw}^2 {ci}->{co} bn={bn}", lambda: D.conv2d(x, w, 3, bn=bn), 2 * n * hw * hw * ci * co * 9))
def gemm(m, n, k, bn=0):
    a = torch.randn(m, k, device=dev).half(); b = (torch.randn(n, k, device=dev) * 0.05).half()
    cases.append((f"gemm {m}x{n}x{k} bn={bn}", lambda: D.gemm(a, b, bn=bn), 2 * m * n * k))
conv(8, 512, 128, 128); conv(8, 256, 256, 256); conv(8, 256, 256, 256, 128); conv(8, 128, 512, 512); conv(8, 128, 512, 512, 128)
conv(24, 64, 320, 320); conv(24, 64, 320, 320, 128); conv(24, 32, 640, 640); conv(24, 16, 1280, 1280); conv(24, 16, 1280, 1280, 128)
conv(24, 8, 2560, 1280); conv(24, 8, 2560, 1280, 64); conv(24, 8, 2560, 1280, 128)
gemm(98304, 320, 320); gemm(98304, 960, 320); gemm(98304, 2560, 320); gemm(98304, 2560, 320, 128); gemm(98304, 320, 1280)
gemm(24576, 5120, 640); gemm(24576, 5120, 640, 128); gemm(6144, 10240, 1280); gemm(6144, 1280, 5120); gemm(1848, 640, 1024)
for name, fn, fl in cases:
    r = []
    for cps in (1, 2):
        lib().dm_tune_gemm(cps)
        ms = timeit(fn)
        r.append(f"cps{cps}: {ms:.3f} ms {fl/ms/1e9:7.1f} TF/s")
    print(f"{name:38s} " + " | ".join(r), flush=True)
