"""Network configurations and weight sources for the dense path.

The reference loads `stabilityai/stable-diffusion-2-1-base` and `zzzyuqing/light-geo-controlnet` through
diffusers (models/guidance/dreammat_guidance.py:88-202).  Here the weights are a flat dict with
diffusers' state-dict key names; they come either from `.safetensors` files (same keys, so the
published checkpoints load unchanged) or -- when no checkpoint is available, as on the benchmark box --
from a seeded random initialisation of the same architecture, generated directly on the device.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass
class UNetConfig:
    """stable-diffusion-2-1-base unet/config.json; `heads` = its attention_head_dim (head_dim = 64)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    heads: Tuple[int, ...] = (5, 10, 20, 20)
    norm_groups: int = 32
    cond_channels: int = 22                       # controlnet_train/diffusers_train_controlnet.py:638
    cond_embed_channels: Tuple[int, ...] = (16, 32, 96, 256)

    @property
    def time_dim(self):
        return self.block_out_channels[0] * 4


@dataclass
class VAEConfig:
    in_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_groups: int = 32
    scaling_factor: float = 0.18215


def load_safetensors(path: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file
    if os.path.isdir(path):
        out: Dict[str, torch.Tensor] = {}
        for fn in sorted(os.listdir(path)):
            if fn.endswith(".safetensors"):
                out.update(load_file(os.path.join(path, fn)))
        if not out:
            raise FileNotFoundError(f"no .safetensors under {path}")
        return out
    return load_file(path)


def unet_config_from_json(path: str, controlnet_json: str = None) -> UNetConfig:
    """diffusers `unet/config.json` (and the ControlNet's `config.json` for the condition-embedding fields) -> UNetConfig.
    Missing files / keys keep the SD-2.1-base defaults (controlnet_train/config.json:2, diffusers_train_controlnet.py:638)."""
    import json
    kw = {}
    if path and os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        heads = j.get("attention_head_dim", None)
        if isinstance(heads, int):
            heads = [heads] * len(j.get("block_out_channels", (320, 640, 1280, 1280)))
        for src, dst in (("in_channels", "in_channels"), ("out_channels", "out_channels"), ("layers_per_block", "layers_per_block"),
                         ("cross_attention_dim", "cross_attention_dim"), ("norm_num_groups", "norm_groups")):
            if src in j:
                kw[dst] = j[src]
        if "block_out_channels" in j:
            kw["block_out_channels"] = tuple(j["block_out_channels"])
        if heads is not None:
            kw["heads"] = tuple(heads)
    if controlnet_json and os.path.exists(controlnet_json):
        with open(controlnet_json) as f:
            j = json.load(f)
        if "conditioning_embedding_out_channels" in j:
            kw["cond_embed_channels"] = tuple(j["conditioning_embedding_out_channels"])
        if "conditioning_channels" in j:
            kw["cond_channels"] = j["conditioning_channels"]
    return UNetConfig(**kw)


def vae_config_from_json(path: str) -> VAEConfig:
    import json
    kw = {}
    if path and os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        for src, dst in (("in_channels", "in_channels"), ("layers_per_block", "layers_per_block"), ("latent_channels", "latent_channels"),
                         ("norm_num_groups", "norm_groups"), ("scaling_factor", "scaling_factor")):
            if src in j:
                kw[dst] = j[src]
        if "block_out_channels" in j:
            kw["block_out_channels"] = tuple(j["block_out_channels"])
    return VAEConfig(**kw)


_LEGACY_VAE_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def normalize_vae_keys(w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Published AutoencoderKL checkpoints (incl. stable-diffusion-2-1-base/vae) store the mid-block attention under the
    deprecated names query / key / value / proj_attn, which diffusers renames at load time
    (AutoencoderKL._convert_deprecated_attention_blocks); some also keep those projections as 1x1 conv weights
    [C, C, 1, 1].  Bring both to the current names / shapes the graph in nets.VAEEncoder expects."""
    out = {}
    for k, v in w.items():
        parts = k.split(".")
        if "attentions" in parts:
            i = parts.index("attentions")
            if len(parts) > i + 2 and parts[i + 2] in _LEGACY_VAE_ATTN:
                parts[i + 2] = _LEGACY_VAE_ATTN[parts[i + 2]]
                k = ".".join(parts)
            if k.endswith((".to_q.weight", ".to_k.weight", ".to_v.weight", ".to_out.0.weight")) and v.dim() == 4:
                v = v.reshape(v.shape[0], v.shape[1])
        out[k] = v
    return out


class _Init:
    """Seeded random tensors in diffusers shapes (variance-preserving scales so a 30-layer fp16 forward
    stays in range).  Device-side generation keeps 1.26 G parameters out of host RAM."""

    def __init__(self, device, seed):
        self.dev = torch.device(device)
        self.g = torch.Generator(device=self.dev).manual_seed(seed)
        self.w: Dict[str, torch.Tensor] = {}

    def randn(self, *shape):
        return torch.randn(*shape, generator=self.g, device=self.dev)

    def conv(self, name, co, ci, k, gain=1.0):
        self.w[name + ".weight"] = self.randn(co, ci, k, k) * (gain / math.sqrt(ci * k * k))
        self.w[name + ".bias"] = self.randn(co) * 0.05

    def lin(self, name, co, ci, bias=True, gain=1.0):
        self.w[name + ".weight"] = self.randn(co, ci) * (gain / math.sqrt(ci))
        if bias:
            self.w[name + ".bias"] = self.randn(co) * 0.05

    def norm(self, name, c):
        self.w[name + ".weight"] = 1.0 + 0.1 * self.randn(c)
        self.w[name + ".bias"] = 0.05 * self.randn(c)

    def resnet(self, p, ci, co, tdim):
        self.norm(p + ".norm1", ci); self.conv(p + ".conv1", co, ci, 3)
        if tdim:
            self.lin(p + ".time_emb_proj", co, tdim)
        self.norm(p + ".norm2", co); self.conv(p + ".conv2", co, co, 3)
        if ci != co:
            self.conv(p + ".conv_shortcut", co, ci, 1)

    def transformer(self, p, c, ctx):
        self.norm(p + ".norm", c); self.lin(p + ".proj_in", c, c)
        b = p + ".transformer_blocks.0"
        for k, kd in (("attn1", c), ("attn2", ctx)):
            self.lin(f"{b}.{k}.to_q", c, c, bias=False); self.lin(f"{b}.{k}.to_k", c, kd, bias=False)
            self.lin(f"{b}.{k}.to_v", c, kd, bias=False); self.lin(f"{b}.{k}.to_out.0", c, c, gain=0.5)
        for k in ("norm1", "norm2", "norm3"):
            self.norm(f"{b}.{k}", c)
        self.lin(f"{b}.ff.net.0.proj", 8 * c, c); self.lin(f"{b}.ff.net.2", c, 4 * c, gain=0.5)
        self.lin(p + ".proj_out", c, c, gain=0.5)

    def encoder_half(self, cfg: UNetConfig):
        ch = cfg.block_out_channels
        self.conv("conv_in", ch[0], cfg.in_channels, 3)
        self.lin("time_embedding.linear_1", cfg.time_dim, ch[0]); self.lin("time_embedding.linear_2", cfg.time_dim, cfg.time_dim)
        ci = ch[0]
        for i, co in enumerate(ch):
            for j in range(cfg.layers_per_block):
                self.resnet(f"down_blocks.{i}.resnets.{j}", ci, co, cfg.time_dim)
                ci = co
                if i < len(ch) - 1:
                    self.transformer(f"down_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
            if i < len(ch) - 1:
                self.conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        self.resnet("mid_block.resnets.0", ch[-1], ch[-1], cfg.time_dim)
        self.transformer("mid_block.attentions.0", ch[-1], cfg.cross_attention_dim)
        self.resnet("mid_block.resnets.1", ch[-1], ch[-1], cfg.time_dim)


def skip_channels(cfg: UNetConfig) -> List[int]:
    ch = cfg.block_out_channels
    out = [ch[0]]
    for i, co in enumerate(ch):
        out += [co] * cfg.layers_per_block
        if i < len(ch) - 1:
            out.append(co)
    return out


def random_unet(cfg: UNetConfig, device="cuda", seed=0) -> Dict[str, torch.Tensor]:
    it = _Init(device, seed)
    it.encoder_half(cfg)
    ch = cfg.block_out_channels
    skips = skip_channels(cfg)
    prev = ch[-1]
    for i, co in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            it.resnet(f"up_blocks.{i}.resnets.{j}", prev + skips.pop(), co, cfg.time_dim)
            prev = co
            if i > 0:
                it.transformer(f"up_blocks.{i}.attentions.{j}", co, cfg.cross_attention_dim)
        if i < len(ch) - 1:
            it.conv(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
    it.norm("conv_norm_out", ch[0]); it.conv("conv_out", cfg.out_channels, ch[0], 3)
    return it.w


def random_controlnet(cfg: UNetConfig, device="cuda", seed=1) -> Dict[str, torch.Tensor]:
    it = _Init(device, seed)
    it.encoder_half(cfg)
    ce = cfg.cond_embed_channels
    it.conv("controlnet_cond_embedding.conv_in", ce[0], cfg.cond_channels, 3)
    k = 0
    for i in range(len(ce) - 1):
        it.conv(f"controlnet_cond_embedding.blocks.{k}", ce[i], ce[i], 3); k += 1
        it.conv(f"controlnet_cond_embedding.blocks.{k}", ce[i + 1], ce[i], 3); k += 1
    it.conv("controlnet_cond_embedding.conv_out", cfg.block_out_channels[0], ce[-1], 3, gain=0.5)
    for i, c in enumerate(skip_channels(cfg)):
        it.conv(f"controlnet_down_blocks.{i}", c, c, 1, gain=0.3)
    it.conv("controlnet_mid_block", cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1, gain=0.3)
    return it.w


def random_vae(cfg: VAEConfig, device="cuda", seed=2) -> Dict[str, torch.Tensor]:
    it = _Init(device, seed)
    ch = cfg.block_out_channels
    it.conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    ci = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            it.resnet(f"encoder.down_blocks.{i}.resnets.{j}", ci, co, 0)
            ci = co
        if i < len(ch) - 1:
            it.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
    it.resnet("encoder.mid_block.resnets.0", ch[-1], ch[-1], 0)
    p = "encoder.mid_block.attentions.0"
    it.norm(p + ".group_norm", ch[-1])
    for k in ("to_q", "to_k", "to_v"):
        it.lin(f"{p}.{k}", ch[-1], ch[-1])
    it.lin(f"{p}.to_out.0", ch[-1], ch[-1], gain=0.5)
    it.resnet("encoder.mid_block.resnets.1", ch[-1], ch[-1], 0)
    it.norm("encoder.conv_norm_out", ch[-1])
    it.conv("encoder.conv_out", 2 * cfg.latent_channels, ch[-1], 3)
    it.conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    return it.w
