"""Attention kernel variants (fp32 vs packed f16x2 exponentials): accuracy vs fp32 torch and graph-replayed timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreammat_b200 import dense_ops as D
from dreammat_b200._cabi import lib
dev = "cuda"
def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
def graph_time(fn, n=10, reps=5):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
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
gen = torch.Generator(device=dev).manual_seed(0)
for (B, h, Nq, Nk, sc) in ((2, 5, 4096, 4096, 1.0), (3, 10, 1024, 1024, 3.0), (2, 20, 64, 77, 1.0), (24, 5, 4096, 4096, 1.0), (24, 10, 1024, 1024, 1.0), (24, 5, 4096, 77, 1.0)):
    C = h * 64
    q = (torch.randn(B, Nq, C, device=dev, generator=gen) * sc).half()
    k = (torch.randn(B, Nk, C, device=dev, generator=gen) * sc).half()
    v = torch.randn(B, Nk, C, device=dev, generator=gen).half()
    ref = None
    if B <= 3:
        qf, kf, vf = (t.float().view(B, -1, h, 64).transpose(1, 2) for t in (q, k, v))
        ref = torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, -1) @ vf
        ref = ref.transpose(1, 2).reshape(B, Nq, C)
    out = []
    for mode in (0, 1, 2):
        lib().dm_tune_attention(mode)
        o = D.attention(q, k, v, h)
        us = graph_time(lambda: D.attention(q, k, v, h)) * 1e3
        fl = 4 * B * h * Nq * Nk * 64
        out.append(f"mode{mode}: {us:8.1f} us {fl/us/1e6:7.1f} TF/s" + (f" rel {rel(o, ref):.2e}" if ref is not None else ""))
    print(f"B{B} h{h} {Nq}x{Nk} scale{sc}: " + " | ".join(out), flush=True)
lib().dm_tune_attention(0)
