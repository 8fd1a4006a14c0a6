"""The oracle's guidance step (a7 / a8 / a9) and the product's schedule logic against vectors produced by executing the
reference's own `StableDiffusionLightGuidance.__call__` / `update_step` end to end -- see tests/golden/make_guidance_golden.py
for what was executed and how the three diffusers modules were stood in for (by the oracle's networks)."""
import os
import types

import pytest
import torch

from oracle import sd as OS

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    return torch.load(os.path.join(HERE, "golden", "guidance_vectors.pt"))


def close(a, b, tol=1e-6):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return a.shape == b.shape and float((a - b).abs().max()) <= tol * (1.0 + float(b.abs().max()))


def test_yaml_schedules_oracle_and_product_host_logic(G):
    """update_step + set_min_max_steps (dreammat_guidance.py:606-628) with the values of configs/dreammat.yaml:54-74: the oracle's C
    and the product's own update_step (host Python, run here on a bare object) reproduce the reference's state at every probe."""
    from dreammat_b200.guidance import StableDiffusionLightGuidance as Prod
    y = G["schedule"]["yaml"]
    cfg = types.SimpleNamespace(**{k: (list(v) if isinstance(v, list) else v) for k, v in y.items()})
    me = types.SimpleNamespace(cfg=cfg, use_controlnet=True, num_train_timesteps=1000)
    me.set_min_max_steps = types.MethodType(Prod.set_min_max_steps, me)
    seen_anneal = set()
    for row in G["schedule"]["trace"]:
        s = row["step"]
        Prod.update_step(me, 0, s)
        got = dict(cond=me.cond_scale, uncond=me.uncond_scale, null=me.null_scale, noise=me.noise_scale, perpneg=me.perpneg_scale)
        for k, v in got.items():
            assert abs(float(v) - row[k]) < 1e-9, (s, k, v, row[k])
            assert abs(float(OS.C(y[k + "_scale"], 0, s)) - row[k]) < 1e-9
        assert (me.min_step, me.max_step) == (row["min_step"], row["max_step"]), (s, me.min_step, me.max_step, row)
        assert list(cfg.condition_scales) == row["condition_scales"], s
        seen_anneal.add(tuple(row["condition_scales"]))
    assert seen_anneal == {(1.0,), (0.8,)}                                   # the probes straddle control_anneal_start_step
    rows = {r["step"]: r for r in G["schedule"]["trace"]}
    # int() truncation in set_min_max_steps: 0.2 + (0.02 - 0.2) * 1.0 = 0.01999999999999999 in binary64 -> min_step 19, not 20
    assert (rows[500]["min_step"], rows[500]["max_step"]) == (200, 800) and (rows[501]["min_step"], rows[501]["max_step"]) == (19, 500)


def test_guidance_call_matches_reference_execution(G):
    c = G["call"]
    gi, go, seen, st = c["in"], c["out"], c["seen"], c["state"]
    ucfg, vcfg = OS.UNetConfig(**gi["unet_cfg"]), OS.VAEConfig(**gi["vae_cfg"])
    wu, wc = OS.random_unet_weights(ucfg, gi["seeds"]["unet"]), OS.random_controlnet_weights(ucfg, gi["seeds"]["controlnet"])
    wv = OS.random_vae_weights(vcfg, gi["seeds"]["vae"])
    B, S = gi["rgb"].shape[0], gi["size"]
    # what the reference handed to the three networks
    x = torch.nn.functional.interpolate(gi["rgb"].permute(0, 3, 1, 2), (S, S), mode="bilinear", align_corners=False) * 2.0 - 1.0
    assert close(x, seen["vae_in"])
    assert seen["controlnet_image_is_condition_map_nchw"] and abs(seen["controlnet_scale"] - 0.8) < 1e-12 and seen["unet_n_down"] == 12
    assert seen["controlnet_t"].dtype == torch.float32 and torch.equal(seen["controlnet_t"], torch.cat([gi["t"]] * 3).float())
    assert st["min_step"] <= int(gi["t"].min()) and int(gi["t"].max()) <= st["max_step"]
    ac = OS.alphas_cumprod()
    # text embeddings of the [text | uncond | null] batch: the product's host mirror of get_text_embeddings
    from dreammat_b200.guidance import PromptProcessorOutput
    tb = gi["tables"]
    pu = PromptProcessorOutput(tb["text"], tb["uncond"], tb["null"], tb["text_vd"], tb["uncond_vd"])
    ctx3 = pu.get_text_embeddings(gi["elevation"], gi["azimuth"], gi["camera_distances"], True, return_null_text_embeddings=True)
    assert close(ctx3, seen["controlnet_ctx"])
    scales = (st["cond"], st["uncond"], st["null"], st["noise"])
    assert abs(scales[1] - (-0.7)) < 1e-9 and abs(scales[2] - (-0.3)) < 1e-9
    rgb = gi["rgb"].clone().requires_grad_(True)
    loss, grad, z, eps = OS.guidance_step(wv, wc, wu, ucfg, vcfg, rgb, gi["condition_map"], ctx3, gi["t"], gi["noise"], gi["vae_eps"],
                                          scales=scales, cond_scale=seen["controlnet_scale"], return_eps=True, resize_to=(S, S))
    zt = ac[gi["t"]].sqrt().view(-1, 1, 1, 1) * z + (1 - ac[gi["t"]]).sqrt().view(-1, 1, 1, 1) * gi["noise"]
    assert close(torch.cat([zt] * 3), seen["controlnet_sample"], 2e-6)
    loss.backward()
    assert close(loss, go["loss_sds"], 1e-5) and close(grad.norm(), go["grad_norm"], 1e-5)
    et, eu, en = eps[0], eps[1], eps[2]
    n = gi["noise"]
    norms = {"uncond_m_noise_norm": eu - n, "text_m_noise_norm": et - n, "text_m_uncond_norm": et - eu, "text_m_null_norm": et - en,
             "null_m_uncond_norm": en - eu, "noise_norm": n, "uncond_norm": eu, "text_norm": et}
    for k, v in norms.items():
        assert close(v.norm(), go[k], 1e-5), k
    assert close(rgb.grad, go["d_rgb"], 1e-5)
    assert float(go["d_rgb"].abs().max()) > 0
