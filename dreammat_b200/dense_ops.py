"""torch-facing wrappers of the dense-path C-ABI entry points (tensor-core GEMM / conv, norms,
attention, elementwise).  Activations are NHWC fp16 (or bf16); accumulation is fp32."""
from __future__ import annotations

import ctypes as C

import torch

from ._cabi import check, lib, ptr, stream_ptr


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rows_per_vec", C.c_int32), ("ld_rowvec", C.c_int32),
                ("residual", C.c_void_p), ("ld_res", C.c_int32), ("res_batch_stride", C.c_int64),
                ("alpha", C.c_float), ("out_scale", C.c_float), ("act", C.c_int32), ("out_f32", C.c_int32)]


ACT = {None: 0, "none": 0, "silu": 1, "gelu": 2}


def _ep(bias=None, rowvec=None, rows_per_vec=1, residual=None, ld_res=0, res_batch_stride=0, alpha=1.0, out_scale=1.0,
        act=None, out_f32=False):
    e = Epilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.rowvec = rowvec.data_ptr() if rowvec is not None else None
    e.rows_per_vec = rows_per_vec
    e.ld_rowvec = rowvec.shape[-1] if rowvec is not None else 0
    e.residual = residual.data_ptr() if residual is not None else None
    e.ld_res = ld_res
    e.res_batch_stride = res_batch_stride
    e.alpha = alpha
    e.out_scale = out_scale
    e.act = ACT[act]
    e.out_f32 = 1 if out_f32 else 0
    return e


def _is_bf16(t):
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float16:
        return 0
    raise TypeError(f"dense path expects fp16/bf16, got {t.dtype}")


def gemm(a, b, bias=None, rowvec=None, rows_per_vec=1, residual=None, alpha=1.0, out_scale=1.0, act=None,
         out=None, out_f32=False, bn=0):
    """a [.., M, K] @ b [.., N, K]^T -> [.., M, N].  Leading dim (if any) is a batch; b may be 2-D (shared)."""
    bf = _is_bf16(a)
    assert b.dtype == a.dtype
    batched = a.dim() == 3
    batch = a.shape[0] if batched else 1
    M, K = a.shape[-2], a.shape[-1]
    N = b.shape[-2]
    assert b.shape[-1] == K
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    lda, ldb = a.stride(-2), b.stride(-2)
    a_bs = a.stride(0) if batched else 0
    b_bs = b.stride(0) if b.dim() == 3 else 0
    if out is None:
        shape = (batch, M, N) if batched else (M, N)
        out = torch.empty(shape, device=a.device, dtype=torch.float32 if out_f32 else a.dtype)
    assert out.stride(-1) == 1
    ldc = out.stride(-2)
    c_bs = out.stride(0) if batched else 0
    ld_res = residual.stride(-2) if residual is not None else 0
    r_bs = residual.stride(0) if (residual is not None and residual.dim() == 3) else 0
    e = _ep(bias, rowvec, rows_per_vec, residual, ld_res, r_bs, alpha, out_scale, act, out.dtype == torch.float32)
    check(lib().dm_gemm(bf, ptr_any(a), lda, a_bs, ptr_any(b), ldb, b_bs, ptr_any(out), ldc, c_bs, M, N, K, batch,
                        C.byref(e), bn, stream_ptr()), "dm_gemm")
    return out


def ptr_any(t):
    return C.c_void_p(t.data_ptr())


def conv2d(x, w, ksize, stride=1, pad=(1, 1), out_hw=None, bias=None, rowvec=None, residual=None, out_scale=1.0,
           act=None, out=None, out_f32=False, bn=0):
    """x [n,H,W,Cin] NHWC (contiguous), w [Cout, k*k*Cin] -> y [n,Ho,Wo,Cout] (or into `out`, whose last-dim
    stride must be 1; its pixel stride gives ldc so it can be a channel slice of a wider buffer)."""
    bf = _is_bf16(x)
    assert x.is_contiguous() and w.is_contiguous()
    n, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == ksize * ksize * Cin
    if out_hw is None:
        pt, pl = pad
        Ho = (H + 2 * pt - ksize) // stride + 1
        Wo = (W + 2 * pl - ksize) // stride + 1
    else:
        Ho, Wo = out_hw
    if out is None:
        out = torch.empty(n, Ho, Wo, Cout, device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    assert out.stride(-1) == 1 and out.shape[:3] == (n, Ho, Wo)
    ldc = out.stride(2)
    assert out.stride(1) == Wo * ldc and out.stride(0) == Ho * Wo * ldc
    ld_res = 0
    if residual is not None:
        ld_res = residual.stride(2)
        assert residual.stride(-1) == 1 and residual.stride(1) == Wo * ld_res and residual.stride(0) == Ho * Wo * ld_res
    e = _ep(bias, rowvec, Ho * Wo, residual, ld_res, 0, 1.0, out_scale, act, out.dtype == torch.float32)
    check(lib().dm_conv2d(bf, ptr_any(x), n, H, W, Cin, ptr_any(w), Cout, ksize, stride, pad[0], pad[1], Ho, Wo,
                          ptr_any(out), ldc, C.byref(e), bn, stream_ptr()), "dm_conv2d")
    return out


def conv_weight_to_gemm(w_oihw: torch.Tensor, cin_pad: int = 0, cout_pad: int = 0, dtype=torch.float16):
    """[Cout,Cin,kh,kw] (diffusers layout) -> [Cout_pad, kh*kw*Cin_pad], tap-major / channel-minor."""
    Cout, Cin, kh, kw = w_oihw.shape
    cin_p = max(cin_pad, Cin)
    cout_p = max(cout_pad, Cout)
    w = torch.zeros(cout_p, kh, kw, cin_p, dtype=torch.float32, device=w_oihw.device)
    w[:Cout, :, :, :Cin] = w_oihw.float().permute(0, 2, 3, 1)
    return w.reshape(cout_p, kh * kw * cin_p).to(dtype).contiguous()
