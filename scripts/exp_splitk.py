"""Split-K + single-occupancy ring on the few-tile layers of a one-view batch: correctness vs fp32 torch and graph-replayed
timing (CPU launch overhead excluded) with the feature on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dreammat_b200 import dense_ops as D
from dreammat_b200._cabi import lib
dev = "cuda"
def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def graph_time(fn, n=20, reps=5):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); 
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps)
g = torch.Generator(device=dev).manual_seed(0)
lib().dm_tune_gemm(21)
print("== correctness (split-K on)", flush=True)
for (M, N, K) in ((192, 1280, 11520), (100, 320, 4096), (768, 64, 2048)):
    a = torch.randn(M, K, device=dev, generator=g).half(); b = (torch.randn(N, K, device=dev, generator=g) * 0.02).half()
    bias = torch.randn(N, device=dev, generator=g).half(); res = torch.randn(M, N, device=dev, generator=g).half()
    for rep in range(2):   # twice: the workspace must come back zeroed
        out = D.gemm(a, b, bias=bias, residual=res, act="silu")
        ref = F.silu(a.float() @ b.float().t() + bias.float()) + res.float()
        print(f"gemm {M}x{N}x{K} pass {rep}: rel {rel(out, ref):.2e}", flush=True)
for (n, hw, ci, co) in ((3, 8, 1280, 1280), (3, 8, 2560, 1280), (1, 16, 128, 64)):
    x = torch.randn(n, hw, hw, ci, device=dev, generator=g).half()
    w = (torch.randn(co, ci, 3, 3, device=dev, generator=g) / 100).half()
    tp = torch.randn(n, co, device=dev, generator=g).half()
    y = D.conv2d(x, D.conv_weight_to_gemm(w), 3, rowvec=tp)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1) + tp.float()[:, :, None, None]
    print(f"conv {n}x{hw}^2 {ci}->{co}: rel {rel(y.float().permute(0, 3, 1, 2), ref):.2e}", flush=True)
print("== timing (graph replay): split-K off | on", flush=True)
cases = []
def conv(n, hw, ci, co):
    x = torch.randn(n, hw, hw, ci, device=dev).half(); w = (torch.randn(co, 9 * ci, device=dev) * 0.02).half()
    cases.append((f"conv {n}x{hw}^2 {ci}->{co}", lambda: D.conv2d(x, w, 3), 2 * n * hw * hw * ci * co * 9))
def gemm(m, n, k):
    a = torch.randn(m, k, device=dev).half(); b = (torch.randn(n, k, device=dev) * 0.05).half()
    cases.append((f"gemm {m}x{n}x{k}", lambda: D.gemm(a, b), 2 * m * n * k))
for B in (3, 24):
    conv(B, 8, 1280, 1280); conv(B, 8, 2560, 1280); conv(B, 16, 1280, 1280); conv(B, 16, 2560, 1280); conv(B, 16, 640, 1280)
    conv(B, 32, 640, 640); conv(B, 32, 1280, 640); conv(B, 64, 320, 320)
    gemm(B * 64, 1280, 1280); gemm(B * 64, 1280, 5120); gemm(B * 64, 10240, 1280); gemm(B * 256, 1280, 1280); gemm(B * 256, 1280, 5120)
    gemm(B * 1024, 640, 640); gemm(B * 1024, 640, 2560); gemm(B * 4096, 320, 320); gemm(B * 77, 2560, 1024)
for name, fn, fl in cases:
    r = []
    for code in (20, 21):
        lib().dm_tune_gemm(code)
        us = graph_time(fn) * 1e3
        r.append(f"{us:8.1f} us {fl/us/1e6:7.1f} TF/s")
    print(f"{name:30s} " + " | ".join(r), flush=True)
