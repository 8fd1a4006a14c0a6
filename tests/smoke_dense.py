"""Dense half of `__graft_entry__.smoke()`: one tiny guidance evaluation (VAE encode with grad, ControlNet + UNet x3,
CSD gradient, VAE backward) on cuda:0, checked against the CPU oracle (the oracle is only the checker here)."""
import torch


def run():
    from oracle import sd as O
    from dreammat_b200 import weights as Wt
    from dreammat_b200.guidance import PromptProcessorOutput, StableDiffusionLightGuidance, _SDSLoss
    Q = lambda x: x.half().float()  # noqa: E731
    ucfg = O.UNetConfig(block_out_channels=(64, 128, 128, 128), heads=(1, 2, 2, 2), cross_attention_dim=64)
    vcfg = O.VAEConfig(block_out_channels=(64, 64, 64, 64))
    wu, wc, wv = (O.round_weights(O.random_unet_weights(ucfg, 0)), O.round_weights(O.random_controlnet_weights(ucfg, 1)),
                  O.round_weights(O.random_vae_weights(vcfg, 2)))
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    B, R = 1, 64
    rgb, cond = torch.rand(B, R, R, 3, generator=g), torch.rand(B, R, R, 22, generator=g)
    vd, uvd, null = torch.randn(4, 77, 64, generator=g), torch.randn(4, 77, 64, generator=g), torch.randn(1, 77, 64, generator=g)
    pu = PromptProcessorOutput(vd[:1], uvd[:1], null, vd, uvd)
    el, az, dist = torch.tensor([10.0]), torch.tensor([20.0]), torch.tensor([3.5])
    t, noise, veps = torch.tensor([400]), torch.randn(B, 4, 8, 8, generator=g), torch.randn(B, 4, 8, 8, generator=g)
    guid = StableDiffusionLightGuidance(dict(use_controlnet=True, control_types=["light"], condition_scales=[1.0], cond_scale=1.05,
                                             uncond_scale=-0.7, null_scale=-0.2), Wt.UNetConfig(**ucfg.__dict__),
                                        Wt.VAEConfig(**vcfg.__dict__), wu, wc, wv, dev)
    x = rgb.to(dev).requires_grad_(True)
    lat = guid.encode_images(x, veps.to(dev))
    ctx3 = pu.get_text_embeddings(el, az, dist, True, return_null_text_embeddings=True)
    grad, dlat, sums = guid.compute_grad_sds(lat, cond.to(dev), ctx3, t.to(dev), noise.to(dev))
    _SDSLoss.apply(lat, dlat, sums[0] / B).backward()
    torch.cuda.synchronize()
    xo = rgb.clone().requires_grad_(True)
    loss_o, grad_o, z_o = O.guidance_step(wv, wc, wu, ucfg, vcfg, xo, cond, Q(ctx3), t, noise, veps, scales=(1.05, -0.7, -0.2, 0.0), q=Q)
    loss_o.backward()
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())  # noqa: E731
    e_z, e_g, e_r = rel(lat.detach(), z_o.detach()), rel(grad, grad_o), rel(x.grad, xo.grad)
    print(f"smoke (dense): latents rel err {e_z:.2e}, sds-grad {e_g:.2e}, d-rgb {e_r:.2e} (fp16 path)")
    assert e_z < 5e-3 and e_g < 5e-2 and e_r < 5e-2, (e_z, e_g, e_r)
