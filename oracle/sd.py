"""Oracle: dense half of the SDS iteration (rows a7-a9 of SURVEY.md section 8).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain-torch restatement of the three diffusers
modules the reference drives from models/guidance/dreammat_guidance.py:
    AutoencoderKL.encode        (:285-292, with autograd)        -> vae_encode
    ControlNetModel.forward     (:205-241 multi_control_forward)  -> controlnet_forward
    UNet2DConditionModel.forward(:262-282 forward_unet)           -> unet_forward
plus compute_grad_sds (:440-497) and the loss tail of __call__ (:584-602).

diffusers itself is an un-vendored, unpinned dependency (requirements.txt:7); the topology below is
its published SD-2.1-base / ControlNet / AutoencoderKL architecture, with
parameter names identical to diffusers' state-dict keys so real checkpoints would load unchanged.
PARITY UNPINNED: no golden outputs exist for this path in the reference tree.

All functions take a flat dict name -> tensor.  `q` is the storage-rounding hook: identity for a
pure fp32 run, `lambda x: x.half().float()` to emulate the reference's fp16-weights /
fp16-activation pipeline (half_precision_weights=True, :56,:92-94) with fp32 accumulation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F

Ident = lambda x: x  # noqa: E731


@dataclass
class UNetConfig:
    """stabilityai/stable-diffusion-2-1-base unet/config.json (attention_head_dim = number of heads)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    heads: Tuple[int, ...] = (5, 10, 20, 20)
    norm_groups: int = 32
    cond_channels: int = 22                      # controlnet conditioning_channels (diffusers_train_controlnet.py:638)
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)

    @property
    def time_dim(self):
        return self.block_out_channels[0] * 4


@dataclass
class VAEConfig:
    """AutoencoderKL encoder of SD-2.1-base vae/config.json."""
    in_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_groups: int = 32
    scaling_factor: float = 0.18215


# ----------------------------------------------------------------------------- building blocks


def conv(w, name, x, stride=1, padding=1):
    return F.conv2d(x, w[name + ".weight"], w.get(name + ".bias"), stride=stride, padding=padding)


def linear(w, name, x):
    return F.linear(x, w[name + ".weight"], w.get(name + ".bias"))


def gn(w, name, x, groups, eps):
    return F.group_norm(x, groups, w[name + ".weight"], w[name + ".bias"], eps)


def timestep_embedding(t, dim):
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def resnet(w, p, x, temb, groups, eps, q):
    """diffusers ResnetBlock2D (output_scale_factor = 1)."""
    h = q(F.silu(gn(w, p + ".norm1", x, groups, eps)))
    h = conv(w, p + ".conv1", h)
    if temb is not None:
        h = h + linear(w, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = q(h)
    h = q(F.silu(gn(w, p + ".norm2", h, groups, eps)))
    h = conv(w, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in w:
        x = conv(w, p + ".conv_shortcut", x, padding=0)
    return q(x + h)


def attention(w, p, x, ctx, heads, q):
    """diffusers Attention (no qkv bias, out bias), head_dim = C / heads, softmax in fp32."""
    B, N, C = x.shape
    src = x if ctx is None else ctx
    qq = q(linear(w, p + ".to_q", x)).view(B, N, heads, C // heads).transpose(1, 2)
    kk = q(linear(w, p + ".to_k", src)).view(B, src.shape[1], heads, C // heads).transpose(1, 2)
    vv = q(linear(w, p + ".to_v", src)).view(B, src.shape[1], heads, C // heads).transpose(1, 2)
    o = F.scaled_dot_product_attention(qq, kk, vv)
    o = q(o.transpose(1, 2).reshape(B, N, C))
    return linear(w, p + ".to_out.0", o)


def transformer(w, p, x, ctx, heads, groups, q):
    """diffusers Transformer2DModel(use_linear_projection=True) with one BasicTransformerBlock."""
    B, C, H, W = x.shape
    res = x
    h = q(gn(w, p + ".norm", x, groups, 1e-6))
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = q(linear(w, p + ".proj_in", h))
    b = p + ".transformer_blocks.0"
    n = q(F.layer_norm(h, (C,), w[b + ".norm1.weight"], w[b + ".norm1.bias"], 1e-5))
    h = q(attention(w, b + ".attn1", n, None, heads, q) + h)
    n = q(F.layer_norm(h, (C,), w[b + ".norm2.weight"], w[b + ".norm2.bias"], 1e-5))
    h = q(attention(w, b + ".attn2", n, ctx, heads, q) + h)
    n = q(F.layer_norm(h, (C,), w[b + ".norm3.weight"], w[b + ".norm3.bias"], 1e-5))
    g = q(linear(w, b + ".ff.net.0.proj", n))
    a, gate = g.chunk(2, dim=-1)
    f = q(a * F.gelu(gate))
    h = q(linear(w, b + ".ff.net.2", f) + h)
    h = linear(w, p + ".proj_out", h)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return q(h + res)


def _time_embed(w, cfg: UNetConfig, t, q):
    temb = q(timestep_embedding(t, cfg.block_out_channels[0]))
    temb = q(F.silu(linear(w, "time_embedding.linear_1", temb)))
    return q(linear(w, "time_embedding.linear_2", temb))


def _down_and_mid(w, cfg: UNetConfig, sample, temb, ctx, q):
    G = cfg.norm_groups
    res = [sample]
    n_blocks = len(cfg.block_out_channels)
    for i in range(n_blocks):
        has_attn = i < n_blocks - 1
        for j in range(cfg.layers_per_block):
            sample = resnet(w, f"down_blocks.{i}.resnets.{j}", sample, temb, G, 1e-5, q)
            if has_attn:
                sample = transformer(w, f"down_blocks.{i}.attentions.{j}", sample, ctx, cfg.heads[i], G, q)
            res.append(sample)
        if i < n_blocks - 1:
            sample = q(conv(w, f"down_blocks.{i}.downsamplers.0.conv", sample, stride=2, padding=1))
            res.append(sample)
    sample = resnet(w, "mid_block.resnets.0", sample, temb, G, 1e-5, q)
    sample = transformer(w, "mid_block.attentions.0", sample, ctx, cfg.heads[-1], G, q)
    sample = resnet(w, "mid_block.resnets.1", sample, temb, G, 1e-5, q)
    return res, sample


def controlnet_forward(w, cfg: UNetConfig, sample, t, ctx, cond, conditioning_scale=1.0, q: Callable = Ident):
    """diffusers ControlNetModel.forward (dreammat_guidance.py:218-229).  sample [N,4,h,w], cond [Nc,22,8h,8w]
    with N = k*Nc: the condition embedding of view b is added to every CFG branch of view b (a0)."""
    temb = _time_embed(w, cfg, t, q)
    sample = conv(w, "conv_in", sample)
    c = q(F.silu(conv(w, "controlnet_cond_embedding.conv_in", cond)))
    nb = 2 * (len(cfg.cond_embed_channels) - 1)
    for i in range(nb):
        c = q(F.silu(conv(w, f"controlnet_cond_embedding.blocks.{i}", c, stride=2 if i % 2 == 1 else 1)))
    c = q(conv(w, "controlnet_cond_embedding.conv_out", c))
    rep = sample.shape[0] // c.shape[0]
    sample = q(sample + c.repeat(rep, 1, 1, 1))
    res, mid = _down_and_mid(w, cfg, sample, temb, ctx, q)
    down = [q(conv(w, f"controlnet_down_blocks.{i}", r, padding=0) * conditioning_scale) for i, r in enumerate(res)]
    mid = q(conv(w, "controlnet_mid_block", mid, padding=0) * conditioning_scale)
    return down, mid


def unet_forward(w, cfg: UNetConfig, sample, t, ctx, down_res=None, mid_res=None, q: Callable = Ident):
    """diffusers UNet2DConditionModel.forward with ControlNet residuals (dreammat_guidance.py:274-282),
    class embedding disabled (:311-317)."""
    G = cfg.norm_groups
    temb = _time_embed(w, cfg, t, q)
    sample = q(conv(w, "conv_in", sample))
    res, sample = _down_and_mid(w, cfg, sample, temb, ctx, q)
    if down_res is not None:
        res = [q(a + b) for a, b in zip(res, down_res)]
    if mid_res is not None:
        sample = q(sample + mid_res)
    n_blocks = len(cfg.block_out_channels)
    for i in range(n_blocks):
        has_attn = i > 0
        for j in range(cfg.layers_per_block + 1):
            skip = res.pop()
            sample = torch.cat([sample, skip], dim=1)
            sample = resnet(w, f"up_blocks.{i}.resnets.{j}", sample, temb, G, 1e-5, q)
            if has_attn:
                sample = transformer(w, f"up_blocks.{i}.attentions.{j}", sample, ctx, cfg.heads[n_blocks - 1 - i], G, q)
        if i < n_blocks - 1:
            sample = F.interpolate(sample, scale_factor=2.0, mode="nearest")
            sample = q(conv(w, f"up_blocks.{i}.upsamplers.0.conv", sample))
    sample = q(F.silu(gn(w, "conv_norm_out", sample, G, 1e-5)))
    return conv(w, "conv_out", sample)


def vae_encode_moments(w, cfg: VAEConfig, x, q: Callable = Ident):
    """diffusers AutoencoderKL.encode -> moments [B, 8, h/8, w/8] (encoder + quant_conv)."""
    G = cfg.norm_groups
    h = q(conv(w, "encoder.conv_in", x))
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet(w, f"encoder.down_blocks.{i}.resnets.{j}", h, None, G, 1e-6, q)
        if i < n - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = q(conv(w, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0))
    h = resnet(w, "encoder.mid_block.resnets.0", h, None, G, 1e-6, q)
    # single-head attention over the 64x64 tokens (head_dim = C), residual connection
    B, C, H, W = h.shape
    p = "encoder.mid_block.attentions.0"
    n_ = q(gn(w, p + ".group_norm", h, G, 1e-6)).permute(0, 2, 3, 1).reshape(B, H * W, C)
    a = attention(w, p, n_, None, 1, q)
    h = q(a.reshape(B, H, W, C).permute(0, 3, 1, 2) + h)
    h = resnet(w, "encoder.mid_block.resnets.1", h, None, G, 1e-6, q)
    h = q(F.silu(gn(w, "encoder.conv_norm_out", h, G, 1e-6)))
    h = q(conv(w, "encoder.conv_out", h))
    return q(conv(w, "quant_conv", h, padding=0))


def vae_sample(moments, eps, scaling, q: Callable = Ident):
    """DiagonalGaussianDistribution.sample() * scaling_factor (dreammat_guidance.py:290-291)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    std = q(torch.exp(0.5 * logvar))
    return q(q(mean + std * q(eps)) * scaling)


# ----------------------------------------------------------------------------- scheduler / CSD


def alphas_cumprod(n=1000, b0=0.00085, b1=0.012):
    """DDIMScheduler(beta_schedule='scaled_linear') as in SD-2.1-base scheduler_config.json."""
    betas = torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def C(value, epoch, global_step):
    """utils/misc.py:65-86."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    s0, v0, v1, s1 = value
    cur = global_step if isinstance(s1, int) else epoch
    return v0 + (v1 - v0) * max(min(1.0, (cur - s0) / (s1 - s0)), 0.0)


def sds_grad(eps_text, eps_uncond, eps_null, noise, t, ac, c, u, n, s):
    """dreammat_guidance.py:475-481 + :584 nan_to_num."""
    wgt = (1 - ac[t]).view(-1, 1, 1, 1)
    return torch.nan_to_num(wgt * (c * eps_text + u * eps_uncond + n * eps_null + s * noise))


def guidance_step(wv, wc, wu, ucfg: UNetConfig, vcfg: VAEConfig, rgb_bhwc, cond_bhwc, ctx3, t, noise, vae_eps,
                  scales=(1.05, -1.0, 0.0, 0.0), cond_scale=1.0, q: Callable = Ident, return_eps: bool = False,
                  resize_to=None):
    """StableDiffusionLightGuidance.__call__ (:536-602) for explicit randomness (appendix B #7-#9).
    ctx3 [3B,77,D] ordered [text | uncond | null].  Returns (loss_sds, grad, latents[, eps [3,B,4,h,w]]).
    `resize_to=(H, W)`: get_latents' rule (:507-513) -- a render whose height differs from cfg.height is resized to
    (cfg.width, cfg.height) = 512 x 512 with bilinear / align_corners=False before the VAE (None: the oracle is also used at
    reduced sizes, where the caller's render IS the VAE input).  Pinned against the reference's own __call__ by
    tests/golden/make_guidance_golden.py."""
    B = rgb_bhwc.shape[0]
    x = rgb_bhwc.permute(0, 3, 1, 2)
    if resize_to is not None and x.shape[2] != resize_to[0]:
        x = F.interpolate(x, tuple(resize_to), mode="bilinear", align_corners=False)
    mom = vae_encode_moments(wv, vcfg, q(x * 2.0 - 1.0), q)
    z = vae_sample(mom, vae_eps, vcfg.scaling_factor, q)
    ac = alphas_cumprod()
    with torch.no_grad():
        zt = ac[t].sqrt().view(-1, 1, 1, 1) * z + (1 - ac[t]).sqrt().view(-1, 1, 1, 1) * noise
        z3 = q(torch.cat([zt] * 3))
        t3 = torch.cat([t] * 3)
        down, mid = controlnet_forward(wc, ucfg, z3, t3, ctx3, q(cond_bhwc.permute(0, 3, 1, 2)), cond_scale, q)
        e = unet_forward(wu, ucfg, z3, t3, ctx3, down, mid, q)
        et, eu, en = e.chunk(3)
        grad = sds_grad(et, eu, en, noise, t, ac, *scales)
    target = (z - grad).detach()
    loss = 0.5 * F.mse_loss(z, target, reduction="sum") / B
    if return_eps:
        return loss, grad, z, e.view(3, B, *e.shape[1:])
    return loss, grad, z


# ----------------------------------------------------------------------------- random weights


def _rand_conv(g, co, ci, k, gain=1.0):
    return torch.randn(co, ci, k, k, generator=g) * (gain / math.sqrt(ci * k * k)), torch.randn(co, generator=g) * 0.05


def _rand_lin(g, co, ci, bias=True, gain=1.0):
    return torch.randn(co, ci, generator=g) * (gain / math.sqrt(ci)), (torch.randn(co, generator=g) * 0.05 if bias else None)


def _put(w, name, wb):
    w[name + ".weight"] = wb[0]
    if wb[1] is not None:
        w[name + ".bias"] = wb[1]


def _norm(g, w, name, c):
    w[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
    w[name + ".bias"] = 0.05 * torch.randn(c, generator=g)


def _rand_resnet(g, w, p, ci, co, tdim):
    _norm(g, w, p + ".norm1", ci)
    _put(w, p + ".conv1", _rand_conv(g, co, ci, 3))
    if tdim:
        _put(w, p + ".time_emb_proj", _rand_lin(g, co, tdim))
    _norm(g, w, p + ".norm2", co)
    _put(w, p + ".conv2", _rand_conv(g, co, co, 3))
    if ci != co:
        _put(w, p + ".conv_shortcut", _rand_conv(g, co, ci, 1))


def _rand_transformer(g, w, p, c, ctx_dim):
    _norm(g, w, p + ".norm", c)
    _put(w, p + ".proj_in", _rand_lin(g, c, c))
    b = p + ".transformer_blocks.0"
    for k, kd in (("attn1", c), ("attn2", ctx_dim)):
        _put(w, f"{b}.{k}.to_q", _rand_lin(g, c, c, bias=False))
        _put(w, f"{b}.{k}.to_k", _rand_lin(g, c, kd, bias=False))
        _put(w, f"{b}.{k}.to_v", _rand_lin(g, c, kd, bias=False))
        _put(w, f"{b}.{k}.to_out.0", _rand_lin(g, c, c, gain=0.5))
    for k in ("norm1", "norm2", "norm3"):
        _norm(g, w, f"{b}.{k}", c)
    _put(w, f"{b}.ff.net.0.proj", _rand_lin(g, 8 * c, c))
    _put(w, f"{b}.ff.net.2", _rand_lin(g, c, 4 * c, gain=0.5))
    _put(w, p + ".proj_out", _rand_lin(g, c, c, gain=0.5))


def _rand_encoder_half(g, w, cfg: UNetConfig):
    ch = cfg.block_out_channels
    _put(w, "conv_in", _rand_conv(g, ch[0], cfg.in_channels, 3))
    _put(w, "time_embedding.linear_1", _rand_lin(g, cfg.time_dim, ch[0]))
    _put(w, "time_embedding.linear_2", _rand_lin(g, cfg.time_dim, cfg.time_dim))
    ci = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _rand_resnet(g, w, f"down_blocks.{i}.resnets.{j}", ci, co, cfg.time_dim)
            ci = co
            if i < len(ch) - 1:
                _rand_transformer(g, w, f"down_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
        if i < len(ch) - 1:
            _put(w, f"down_blocks.{i}.downsamplers.0.conv", _rand_conv(g, co, co, 3))
    _rand_resnet(g, w, "mid_block.resnets.0", ch[-1], ch[-1], cfg.time_dim)
    _rand_transformer(g, w, "mid_block.attentions.0", ch[-1], cfg.cross_attention_dim)
    _rand_resnet(g, w, "mid_block.resnets.1", ch[-1], ch[-1], cfg.time_dim)


def skip_channels(cfg: UNetConfig) -> List[int]:
    ch = cfg.block_out_channels
    out = [ch[0]]
    for i, co in enumerate(ch):
        out += [co] * cfg.layers_per_block
        if i < len(ch) - 1:
            out.append(co)
    return out


def random_unet_weights(cfg: UNetConfig, seed=0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    _rand_encoder_half(g, w, cfg)
    ch = cfg.block_out_channels
    skips = skip_channels(cfg)
    rev = list(reversed(ch))
    prev = ch[-1]
    for i, co in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            sk = skips.pop()
            _rand_resnet(g, w, f"up_blocks.{i}.resnets.{j}", prev + sk, co, cfg.time_dim)
            prev = co
            if i > 0:
                _rand_transformer(g, w, f"up_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
        if i < len(ch) - 1:
            _put(w, f"up_blocks.{i}.upsamplers.0.conv", _rand_conv(g, co, co, 3))
    _norm(g, w, "conv_norm_out", ch[0])
    _put(w, "conv_out", _rand_conv(g, cfg.out_channels, ch[0], 3))
    return w


def random_controlnet_weights(cfg: UNetConfig, seed=1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    _rand_encoder_half(g, w, cfg)
    ce = cfg.cond_embed_channels
    _put(w, "controlnet_cond_embedding.conv_in", _rand_conv(g, ce[0], cfg.cond_channels, 3))
    k = 0
    for i in range(len(ce) - 1):
        _put(w, f"controlnet_cond_embedding.blocks.{k}", _rand_conv(g, ce[i], ce[i], 3)); k += 1
        _put(w, f"controlnet_cond_embedding.blocks.{k}", _rand_conv(g, ce[i + 1], ce[i], 3)); k += 1
    # zero-initialised in a fresh ControlNet; random here so the parity test exercises them
    _put(w, "controlnet_cond_embedding.conv_out", _rand_conv(g, cfg.block_out_channels[0], ce[-1], 3, gain=0.5))
    for i, c in enumerate(skip_channels(cfg)):
        _put(w, f"controlnet_down_blocks.{i}", _rand_conv(g, c, c, 1, gain=0.3))
    _put(w, "controlnet_mid_block", _rand_conv(g, cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1, gain=0.3))
    return w


def random_vae_weights(cfg: VAEConfig, seed=2) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    ch = cfg.block_out_channels
    _put(w, "encoder.conv_in", _rand_conv(g, ch[0], cfg.in_channels, 3))
    ci = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _rand_resnet(g, w, f"encoder.down_blocks.{i}.resnets.{j}", ci, co, 0)
            ci = co
        if i < len(ch) - 1:
            _put(w, f"encoder.down_blocks.{i}.downsamplers.0.conv", _rand_conv(g, co, co, 3))
    _rand_resnet(g, w, "encoder.mid_block.resnets.0", ch[-1], ch[-1], 0)
    p = "encoder.mid_block.attentions.0"
    _norm(g, w, p + ".group_norm", ch[-1])
    for k in ("to_q", "to_k", "to_v"):
        _put(w, f"{p}.{k}", _rand_lin(g, ch[-1], ch[-1]))
    _put(w, f"{p}.to_out.0", _rand_lin(g, ch[-1], ch[-1], gain=0.5))
    _rand_resnet(g, w, "encoder.mid_block.resnets.1", ch[-1], ch[-1], 0)
    _norm(g, w, "encoder.conv_norm_out", ch[-1])
    _put(w, "encoder.conv_out", _rand_conv(g, 2 * cfg.latent_channels, ch[-1], 3))
    _put(w, "quant_conv", _rand_conv(g, 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1))
    return w


def round_weights(w, dtype=torch.float16):
    return {k: v.to(dtype).float() for k, v in w.items()}
