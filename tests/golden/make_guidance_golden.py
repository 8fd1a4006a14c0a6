"""Golden vectors for the guidance (a7 / a8 / a9), made by EXECUTING the reference's own `StableDiffusionLightGuidance`
methods end to end: `update_step`, `set_min_max_steps`, `__call__`, `get_latents`, `encode_images`, `prepare_image_cond`,
`compute_grad_sds`, `compute_without_perpneg`, `multi_control_forward`, `forward_controlnet`-style plumbing, `forward_unet`,
`disable_unet_class_embedding` (models/guidance/dreammat_guidance.py:205-316,388-640), plus `utils/misc.C` and the prompt
processor's `get_text_embeddings` (as in make_golden.py).

The three diffusers modules are absent here; the ORACLE's networks (oracle/sd.py, small widths, fp32, seeded weights) stand in
for them behind diffusers' call signatures:

  self.vae.encode(x).latent_dist.sample()   -> oracle VAE moments; sample() draws torch.randn itself (the draw order is pinned)
  self.controlnet.nets[0](sample, t, ctx, image, scale, ..., return_dict=False) -> oracle ControlNet
  unet(latents, t, encoder_hidden_states=..., down_block_additional_residuals=..., mid_block_additional_residual=...).sample
  self.scheduler.add_noise                   -> sqrt(a_t) x + sqrt(1 - a_t) n on the SD scaled-linear schedule (published)

so what the vectors pin is everything the REFERENCE wrote between those calls: the yaml's schedules over the training steps
(scales, min / max timestep with its int() truncation, the condition-scale anneal), the resize of a non-512 render, `x * 2 - 1`,
the scaling factor, the order of the three random draws (posterior, timestep, noise), the [text | uncond | null] batch, the
condition image layout and scale, the residual plumbing, the CSD combination, loss and the gradient that reaches the render.

Run from the repo root where /root/reference exists:  python tests/golden/make_guidance_golden.py -> guidance_vectors.pt
"""
import contextlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sd as OS                                                   # noqa: E402
from tests.golden.make_golden import REF, Fake, base_ns, lift, lift_class     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "guidance_vectors.pt")
# configs/dreammat.yaml:54-74
YAML = dict(cond_scale=1.05, uncond_scale=[0, -1.0, -0.5, 2000], null_scale=[0, 0.0, -0.5, 2000], noise_scale=0.0, perpneg_scale=0.0,
            min_step_percent=[500, 0.2, 0.02, 501], max_step_percent=[500, 0.8, 0.5, 501], control_anneal_start_step=700,
            condition_scales=[1.0], condition_scales_anneal=[0.8])
UCFG = dict(block_out_channels=(32, 64, 64, 64), heads=(1, 2, 2, 2), cross_attention_dim=32, cond_embed_channels=(8, 16, 16, 32))
VCFG = dict(block_out_channels=(32, 32, 32, 32))
SEEDS = dict(unet=10, controlnet=11, vae=12)
SIZE = 64                # stands for cfg.height = cfg.width = 512


def main():
    ng = base_ns()
    nc = base_ns(); nc["config_to_primitive"] = lambda v: list(v)
    lift("utils/misc.py", ["C"], nc)
    ng["C"] = nc["C"]
    ng["threestudio"] = Fake(info=lambda *a, **k: None)
    names = ["update_step", "set_min_max_steps", "__call__", "get_latents", "encode_images", "prepare_image_cond", "compute_grad_sds",
             "compute_without_perpneg", "multi_control_forward", "forward_unet", "disable_unet_class_embedding"]
    lift("models/guidance/dreammat_guidance.py", names, ng, cls="StableDiffusionLightGuidance")
    ng["disable_unet_class_embedding"] = contextlib.contextmanager(ng["disable_unet_class_embedding"])   # its own decorator (:310)

    def fresh_cfg():
        return Fake(**{k: (list(v) if isinstance(v, list) else v) for k, v in YAML.items()}, control_types=["light"], height=SIZE,
                    width=SIZE, view_dependent_prompting=True, grad_clip_val=None, grad_normalize=False)

    # ---------------------------------------------------------------- (A) the schedules of dreammat.yaml over the run
    gd = Fake(cfg=fresh_cfg(), use_controlnet=True, num_train_timesteps=1000).bind(ng, ["update_step", "set_min_max_steps"])
    trace = []
    for step in (0, 1, 250, 499, 500, 501, 699, 700, 701, 1000, 1500, 1999, 2000, 2001, 3000):
        gd.update_step(0, step)
        trace.append(dict(step=step, cond=float(gd.cond_scale), uncond=float(gd.uncond_scale), null=float(gd.null_scale),
                          noise=float(gd.noise_scale), perpneg=float(gd.perpneg_scale), min_step=int(gd.min_step), max_step=int(gd.max_step),
                          condition_scales=list(gd.cfg.condition_scales)))

    # ---------------------------------------------------------------- (B) one full __call__
    ucfg, vcfg = OS.UNetConfig(**UCFG), OS.VAEConfig(**VCFG)
    wu, wc, wv = OS.random_unet_weights(ucfg, SEEDS["unet"]), OS.random_controlnet_weights(ucfg, SEEDS["controlnet"]), OS.random_vae_weights(vcfg, SEEDS["vae"])
    calls = []

    class Posterior:
        def __init__(self, mom):
            self.mom = mom

        def sample(self):
            eps = torch.randn(self.mom.shape[0], self.mom.shape[1] // 2, *self.mom.shape[2:])
            calls.append(("vae_eps", eps.clone()))
            return OS.vae_sample(self.mom, eps, 1.0)

    class VAE:
        config = Fake(scaling_factor=vcfg.scaling_factor)

        def encode(self, x):
            calls.append(("vae_in", x.detach().clone()))
            return Fake(latent_dist=Posterior(OS.vae_encode_moments(wv, vcfg, x)))

    def controlnet(sample, timestep, encoder_hidden_states, image, scale, class_labels, timestep_cond, attention_mask,
                   cross_attention_kwargs, return_dict=False):
        assert class_labels is None and timestep_cond is None and attention_mask is None and not return_dict
        calls.append(("controlnet", dict(sample=sample.clone(), t=timestep.clone(), ctx=encoder_hidden_states.clone(), image=image.clone(), scale=scale)))
        return OS.controlnet_forward(wc, ucfg, sample, timestep, encoder_hidden_states, image, scale)

    class UNet:
        class_embedding = "sentinel"

        def __call__(self, latents, t, encoder_hidden_states=None, class_labels=None, cross_attention_kwargs=None,
                     down_block_additional_residuals=None, mid_block_additional_residual=None):
            assert self.class_embedding is None and class_labels is None          # :311-317 disables it around the call
            calls.append(("unet", dict(t=t.clone(), n_down=len(down_block_additional_residuals))))
            return Fake(sample=OS.unet_forward(wu, ucfg, latents, t, encoder_hidden_states, down_block_additional_residuals,
                                               mid_block_additional_residual))
    alphas = OS.alphas_cumprod()

    class Sched:
        def add_noise(self, x, n, t):
            a = alphas[t].view(-1, 1, 1, 1)
            return a.sqrt() * x + (1 - a).sqrt() * n
    # prompt processor output: the reference's own selection code over small random tables
    npp = base_ns()
    lift("models/prompt_processors/base.py", ["shift_azimuth_deg"], npp)
    lift_class("models/prompt_processors/base.py", "DirectionConfig", npp, keep_fields=True)
    lift("models/prompt_processors/base.py", ["configure"], npp, cls="PromptProcessor")
    lift("models/prompt_processors/base.py", ["get_text_embeddings"], npp, cls="PromptProcessorOutput")
    pp = Fake(cfg=Fake(view_dependent_prompt_front=False, front_threshold=45.0, back_threshold=45.0, overhead_threshold=60.0)).bind(npp, ["configure"])
    try:
        pp.configure()
    except (FileNotFoundError, OSError):
        pass
    g = torch.Generator().manual_seed(77)
    NT, D = 7, ucfg.cross_attention_dim
    tables = {"text": torch.randn(1, NT, D, generator=g), "uncond": torch.randn(1, NT, D, generator=g), "null": torch.randn(1, NT, D, generator=g),
              "text_vd": torch.randn(4, NT, D, generator=g), "uncond_vd": torch.randn(4, NT, D, generator=g)}
    po = Fake(text_embeddings=tables["text"], uncond_text_embeddings=tables["uncond"], null_text_embeddings=tables["null"],
              text_embeddings_vd=tables["text_vd"], uncond_text_embeddings_vd=tables["uncond_vd"], directions=pp.directions,
              direction2idx=pp.direction2idx, use_perp_neg=False).bind(npp, ["get_text_embeddings"])
    unet = UNet()
    gd = Fake(cfg=fresh_cfg(), use_controlnet=True, num_train_timesteps=1000, device="cpu", weights_dtype=torch.float32, vae=VAE(),
              unet=unet, controlnet=Fake(nets=[controlnet]), scheduler=Sched(), alphas=alphas).bind(ng, names)
    STEP = 1200
    gd.update_step(0, STEP)                   # uncond -0.7, null -0.3, condition scale annealed to 0.8, t in [19, 500]
    B, R = 2, 40
    rgb = torch.rand(B, R, R, 3, generator=g).requires_grad_(True)          # a 40 x 40 render: get_latents resizes it to SIZE
    cond = torch.rand(B, SIZE, SIZE, 22, generator=g)
    el, az, dist = torch.tensor([10.0, 70.0]), torch.tensor([100.0, -30.0]), torch.tensor([3.0, 3.2])
    SEED = 2024
    drawn = []
    real_randn_like = torch.randn_like

    def recording_randn_like(x, *a, **k):     # :461 -- recorded, not replayed: `latents` arrives channels-last from the stand-in
        r = real_randn_like(x, *a, **k)       # VAE and randn_like fills in memory order, so its values are layout-dependent
        drawn.append(r.clone())
        return r
    torch.randn_like = recording_randn_like
    try:
        torch.manual_seed(SEED)
        out = gd.__call__(rgb, po, el, az, dist, torch.tensor([0, 3]), rgb_as_latents=False, condition_map=cond)
    finally:
        torch.randn_like = real_randn_like
    out["loss_sds"].backward()
    torch.manual_seed(SEED)                   # the draws of the call, in order: :290 posterior, :453 timestep (then :461 noise)
    vae_eps = torch.randn(B, 4, SIZE // 8, SIZE // 8)
    t = torch.randint(gd.min_step, gd.max_step + 1, [B], dtype=torch.long)
    assert len(drawn) == 1
    noise = drawn[0].contiguous()
    seq = [c[0] for c in calls]
    assert seq == ["vae_in", "vae_eps", "controlnet", "unet"], seq
    assert torch.equal(calls[1][1], vae_eps) and unet.class_embedding == "sentinel"
    cn = calls[2][1]
    image_is_permuted_map = bool(torch.equal(cn["image"], cond.permute(0, 3, 1, 2)))     # prepare_image_cond('light'): NHWC -> NCHW, no resize
    G = {"schedule": {"yaml": YAML, "trace": trace},
         "call": {"in": {"rgb": rgb.detach(), "condition_map": cond, "elevation": el, "azimuth": az, "camera_distances": dist, "tables": tables,
                         "global_step": STEP, "size": SIZE, "unet_cfg": UCFG, "vae_cfg": VCFG, "seeds": SEEDS, "vae_eps": vae_eps, "t": t,
                         "noise": noise},
                  "seen": {"vae_in": calls[0][1], "controlnet_sample": cn["sample"], "controlnet_t": cn["t"], "controlnet_ctx": cn["ctx"],
                           "controlnet_image_is_condition_map_nchw": image_is_permuted_map, "controlnet_scale": float(cn["scale"]), "unet_t": calls[3][1]["t"],
                           "unet_n_down": calls[3][1]["n_down"]},
                  "state": {"cond": float(gd.cond_scale), "uncond": float(gd.uncond_scale), "null": float(gd.null_scale), "noise": float(gd.noise_scale),
                            "min_step": int(gd.min_step), "max_step": int(gd.max_step)},
                  "out": {k: v.detach() for k, v in out.items()} | {"d_rgb": rgb.grad.clone()}}}
    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; t =", t.tolist(), "loss_sds", float(out["loss_sds"].detach()), "keys", sorted(out))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    main()
