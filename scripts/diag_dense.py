"""GPU diagnostic for the streaming dense kernels and the fused attention kernel (vs torch fp32 on the GPU)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dreammat_b200 import dense_ops as D

torch.manual_seed(0)
dev = "cuda"
H16 = torch.float16


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def case(name, fn, tol=3e-3):
    try:
        r = fn()
        torch.cuda.synchronize()
        print(f"{name:60s} rel_err={r:.3e} {'OK' if r < tol else 'BAD'}", flush=True)
    except Exception as e:
        print(f"{name:60s} EXC {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def attn_case(B, heads, Nq, Nk, dtype=H16, fused_qkv=False):
    def f():
        Cc = heads * 64
        if fused_qkv:
            qkv = (torch.randn(B, Nq, 3 * Cc, device=dev)).to(dtype)
            q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
        else:
            q = torch.randn(B, Nq, Cc, device=dev).to(dtype)
            k = torch.randn(B, Nk, Cc, device=dev).to(dtype)
            v = torch.randn(B, Nk, Cc, device=dev).to(dtype)
        o = D.attention(q, k, v, heads)
        qf = q.float().view(B, Nq, heads, 64).transpose(1, 2)
        kf = k.float().view(B, -1, heads, 64).transpose(1, 2)
        vf = v.float().view(B, -1, heads, 64).transpose(1, 2)
        ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Nq, Cc)
        return rel(o.float(), ref)
    return f


def gn_case(n, Hh, W, Cc, silu, dtype=H16):
    def f():
        x = (torch.randn(n, Hh, W, Cc, device=dev) * 1.5 + 0.3).to(dtype)
        g = (torch.randn(Cc, device=dev) * 0.5 + 1).to(dtype); b = (torch.randn(Cc, device=dev) * 0.2).to(dtype)
        y, st = D.groupnorm(x, g, b, silu=silu)
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        ref = F.group_norm(xr, 32, g.float(), b.float(), 1e-5)
        if silu: ref = F.silu(ref)
        e1 = rel(y.float().permute(0, 3, 1, 2), ref)
        dz = torch.randn(n, Hh, W, Cc, device=dev).to(dtype)
        ref.backward(dz.float().permute(0, 3, 1, 2))
        dx = D.groupnorm_bwd(x, dz, g, b, st, silu=silu)
        e2 = rel(dx.float().permute(0, 3, 1, 2), xr.grad)
        print(f"      gn fwd {e1:.2e} bwd {e2:.2e}")
        return max(e1, e2)
    return f


case("attention self B2 h5 N4096", attn_case(2, 5, 4096, 4096))
case("attention self B2 h5 N4096 fused qkv view", attn_case(2, 5, 4096, 4096, fused_qkv=True))
case("attention self B3 h10 N1024", attn_case(3, 10, 1024, 1024))
case("attention self B3 h20 N256", attn_case(3, 20, 256, 256))
case("attention self B3 h20 N64", attn_case(3, 20, 64, 64))
case("attention cross B3 h5 Nq4096 Nk77", attn_case(3, 5, 4096, 77))
case("attention cross B2 h20 Nq64 Nk77", attn_case(2, 20, 64, 77))
case("attention bf16 B1 h5 N1024", attn_case(1, 5, 1024, 1024, dtype=torch.bfloat16), tol=1e-2)
case("groupnorm+silu 2x64x64x320", gn_case(2, 64, 64, 320, True))
case("groupnorm 1x32x32x1280", gn_case(1, 32, 32, 1280, False))
case("groupnorm+silu 1x128x128x128", gn_case(1, 128, 128, 128, True))
case("groupnorm+silu 3x8x8x2560", gn_case(3, 8, 8, 2560, True))
case("groupnorm+silu 2x16x16x512", gn_case(2, 16, 16, 512, True))


def ln():
    x = torch.randn(3, 1024, 640, device=dev).half(); g = torch.randn(640, device=dev).half(); b = torch.randn(640, device=dev).half()
    return rel(D.layernorm(x, g, b).float(), F.layer_norm(x.float(), (640,), g.float(), b.float()))
case("layernorm 3x1024x640", ln)


def gg():
    h = torch.randn(2, 256, 2 * 1280, device=dev).half()
    a, g = h.float().chunk(2, -1)
    return rel(D.geglu(h).float(), a * F.gelu(g))
case("geglu", gg)


def up():
    x = torch.randn(2, 8, 8, 64, device=dev).half()
    e1 = rel(D.upsample2x(x).float().permute(0, 3, 1, 2), F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest"))
    z = D.upsample2x(x, zero_insert=True)
    ref = torch.zeros(2, 16, 16, 64, device=dev); ref[:, ::2, ::2] = x.float()
    return max(e1, rel(z.float(), ref))
case("upsample2x / zero-insert", up)


def ax():
    big = torch.zeros(4, 16, 16, 192, device=dev).half()
    s = torch.randn(4, 16, 16, 64, device=dev).half(); t = torch.randn(4, 16, 16, 64, device=dev).half()
    D.axpby(s.view(-1, 64), 1.0, out=big.view(-1, 192)[:, 64:128])
    e1 = rel(big[..., 64:128].float(), s.float()) + float(big[..., :64].abs().sum()) + float(big[..., 128:].abs().sum())
    r = D.axpby(s.view(-1, 64), 0.5, t.view(-1, 64), 2.0)
    return max(e1, rel(r.float(), 0.5 * s.float().view(-1, 64) + 2 * t.float().view(-1, 64)))
case("axpby2d (slice copy, a*x+b*y)", ax)


def tr():
    x = torch.randn(3, 100, 77, device=dev).half()
    return rel(D.transpose(x).float(), x.float().transpose(1, 2))
case("transpose", tr)


def sm():
    x = torch.randn(64, 4096, device=dev).half() * 4
    Pm = D.softmax_rows(x, 0.3)
    e1 = rel(Pm.float(), torch.softmax(x.float() * 0.3, -1))
    dP = torch.randn(64, 4096, device=dev).half()
    xr = x.float().requires_grad_(True)
    torch.softmax(xr * 0.3, -1).backward(dP.float())
    e2 = rel(D.softmax_bwd(Pm, dP, 0.3).float(), xr.grad)
    print(f"      softmax fwd {e1:.2e} bwd {e2:.2e}")
    return max(e1, e2)
case("softmax rows fwd/bwd", sm, tol=6e-3)


def pc():
    x = torch.rand(2, 16, 16, 3, device=dev)
    y = D.pad_convert(x, 64, 2.0, -1.0)
    e1 = rel(y[..., :3].float(), 2 * x - 1) + float(y[..., 3:].abs().sum())
    w = D.unpad_convert(y, 3, 2.0)
    return max(e1, rel(w, 2 * (2 * x - 1)))
case("pad/unpad convert", pc)


def vs():
    mom = torch.randn(2, 8, 8, 64, device=dev).half(); eps = torch.randn(2, 4, 8, 8, device=dev)
    z = D.vae_sample(mom, eps)
    m = mom.float().permute(0, 3, 1, 2)
    mean, lv = m[:, :4], m[:, 4:8].clamp(-30, 20)
    ref = (mean + torch.exp(0.5 * lv) * eps) * 0.18215
    return rel(z, ref)
case("vae sample", vs)


def an():
    z = torch.randn(2, 4, 8, 8, device=dev); nz = torch.randn(2, 4, 8, 8, device=dev)
    a = torch.tensor([0.9, 0.5], device=dev); b = torch.tensor([0.4, 0.8], device=dev)
    o = D.add_noise(z, nz, a, b)
    ref = (a.view(2, 1, 1, 1) * z + b.view(2, 1, 1, 1) * nz).permute(0, 2, 3, 1)
    return rel(o[:2, ..., :4].float(), ref) + rel(o[4:6, ..., :4].float(), ref) + float(o[..., 4:].abs().sum())
case("add_noise x3", an)


def te():
    t = torch.tensor([10.0, 500.0, 980.0], device=dev)
    o = D.timestep_embedding(t, 320)
    half = 160
    f = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, device=dev) / half)
    a = t[:, None] * f[None]
    return rel(o.float(), torch.cat([torch.cos(a), torch.sin(a)], -1))
case("timestep embedding", te)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
try:
    q = torch.randn(24, 4096, 960, device=dev).half()
    ms = timeit(lambda: D.attention(q[..., :320], q[..., 320:640], q[..., 640:], 5))
    print(f"attention 24x5x4096x4096x64: {ms:.3f} ms  {4*24*5*4096*4096*64/ms/1e9:.1f} TFLOP/s")
    x = torch.randn(8, 512, 512, 128, device=dev).half(); g = torch.ones(128, device=dev).half(); b = torch.zeros(128, device=dev).half()
    ms = timeit(lambda: D.groupnorm(x, g, b, silu=True))
    print(f"groupnorm+silu 8x512x512x128: {ms:.3f} ms  {3*x.numel()*2/ms/1e6:.1f} GB/s (2 reads + 1 write)")
except Exception as e:
    print("timing EXC", e)
