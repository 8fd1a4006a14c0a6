"""CPU checks that pin the dense-side oracle's topology (no GPU)."""
import torch

from oracle import sd as O


def test_parameter_counts_match_published_sd21_base():
    """The only public numbers available offline: UNet 865,910,724 parameters (SD-2.x UNet), VAE encoder
    + quant_conv 34,163,664.  Shapes only -- built on the meta device."""
    with torch.device("meta"):
        wu = O.random_unet_weights(O.UNetConfig())
        wv = O.random_vae_weights(O.VAEConfig())
        wc = O.random_controlnet_weights(O.UNetConfig())
    assert sum(v.numel() for v in wu.values()) == 865910724
    assert sum(v.numel() for v in wv.values()) == 34163664
    # ControlNetModel.from_unet(unet, conditioning_channels=22): encoder half + cond embedding + 13 zero convs
    assert wc["controlnet_cond_embedding.conv_in.weight"].shape == (16, 22, 3, 3)
    assert len([k for k in wc if k.startswith("controlnet_down_blocks.") and k.endswith(".weight")]) == 12
    assert O.skip_channels(O.UNetConfig()) == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]


def test_scheduler_and_schedules():
    ac = O.alphas_cumprod()
    assert ac.shape == (1000,) and abs(float(ac[0]) - 0.99915) < 1e-5 and float(ac[-1]) < 0.01
    # configs/dreammat.yaml:63-68
    assert O.C([0, -1.0, -0.5, 2000], 0, 1000) == -0.75
    assert O.C([500, 0.2, 0.02, 501], 0, 500) == 0.2 and abs(O.C([500, 0.2, 0.02, 501], 0, 501) - 0.02) < 1e-12
    assert O.C(1.05, 0, 10) == 1.05


def test_tiny_guidance_step_runs_and_differentiates():
    ucfg = O.UNetConfig(block_out_channels=(64, 64, 64, 64), heads=(1, 1, 1, 1), cross_attention_dim=64)
    vcfg = O.VAEConfig(block_out_channels=(32, 32, 32, 32))
    wu, wc, wv = O.random_unet_weights(ucfg), O.random_controlnet_weights(ucfg), O.random_vae_weights(vcfg)
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(1, 64, 64, 3, generator=g).requires_grad_(True)
    loss, grad, z = O.guidance_step(wv, wc, wu, ucfg, vcfg, rgb, torch.rand(1, 64, 64, 22, generator=g),
                                    torch.randn(3, 77, 64, generator=g), torch.tensor([400]),
                                    torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g))
    loss.backward()
    # d loss / d z = grad / B  (reparameterisation, dreammat_guidance.py:590-594)
    assert torch.isfinite(rgb.grad).all() and float(rgb.grad.abs().sum()) > 0
    assert abs(float(loss) - 0.5 * float((grad ** 2).sum())) < 1e-4 * float(loss)
