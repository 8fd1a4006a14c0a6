"""GPU: the drop-in boundary end to end.  `threestudio.find("dreammat-system")(cfg)` with the values of
configs/dreammat.yaml:28-115 builds geometry / material / renderer through the registry, `on_fit_start` builds the guidance
(weights read from local diffusers-format directories: *.safetensors + config.json) and the prompt processor, and the
reference's training loop body (Lightning automatic optimisation: update walk -> training_step -> zero_grad -> backward ->
optimizer.step, systems/base.py + systems/dreammat.py:57-86) runs against it.  The package under `threestudio` is the stub
of tests/stub_threestudio (pinned against the reference in tests/test_plugin_registry.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from tests.test_plugin_registry import FIVE, STUB, YAML_SYSTEM, _write_obj

pytestmark = pytest.mark.gpu


def _write_envmaps(root):
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    from dreammat_b200.scene import synthetic_envmap
    for i in range(1, 6):
        d = os.path.join(root, f"map{i}")
        os.makedirs(d, exist_ok=True)
        img = synthetic_envmap(64, 128, seed=i).numpy().astype(np.float32)
        assert cv2.imwrite(os.path.join(d, f"map{i}.exr"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))


def _write_checkpoints(root):
    """tiny diffusers-format checkpoints: <root>/stabilityai/stable-diffusion-2-1-base/{unet,vae}/ + <root>/controlnet/"""
    from safetensors.torch import save_file
    from dreammat_b200 import weights as W
    ucfg = W.UNetConfig(block_out_channels=(64, 128, 128, 128), heads=(1, 2, 2, 2), cross_attention_dim=64)
    vcfg = W.VAEConfig(block_out_channels=(64, 64, 64, 64))
    base = os.path.join(root, "stabilityai", "stable-diffusion-2-1-base")
    for sub, w, cj in (("unet", W.random_unet(ucfg, "cpu", 0), {"attention_head_dim": list(ucfg.heads), "block_out_channels": list(ucfg.block_out_channels),
                                                                "cross_attention_dim": 64, "layers_per_block": 2, "norm_num_groups": 32}),
                       ("vae", W.random_vae(vcfg, "cpu", 2), {"block_out_channels": list(vcfg.block_out_channels), "latent_channels": 4,
                                                              "scaling_factor": 0.18215, "norm_num_groups": 32})):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
        if sub == "vae":   # published VAE checkpoints carry the deprecated attention names
            ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
            w2 = {}
            for k, v in w.items():
                for a, b in ren.items():
                    k = k.replace(f"attentions.0.{a}.", f"attentions.0.{b}.")
                w2[k] = v
            w = w2
        save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(base, sub, "diffusion_pytorch_model.safetensors"))
        with open(os.path.join(base, sub, "config.json"), "w") as f:
            json.dump(cj, f)
    cn = os.path.join(root, "controlnet")
    os.makedirs(cn, exist_ok=True)
    save_file({k: v.contiguous() for k, v in W.random_controlnet(ucfg, "cpu", 1).items()}, os.path.join(cn, "diffusion_pytorch_model.safetensors"))
    with open(os.path.join(cn, "config.json"), "w") as f:
        json.dump({"conditioning_embedding_out_channels": [16, 32, 96, 256], "conditioning_channels": 22}, f)
    return ucfg


def test_one_training_step_through_the_registry(tmp_path):
    for m in [k for k in sys.modules if k == "threestudio" or k.startswith("threestudio.") or k == "dreammat_b200.threestudio_plugin"]:
        del sys.modules[m]
    sys.path.insert(0, STUB)
    try:
        import threestudio
        from threestudio.utils.base import BaseModule, BaseObject
        import dreammat_b200.threestudio_plugin  # noqa: F401
        from dreammat_b200.guidance import PromptProcessorOutput
        from dreammat_b200.scene import DataConfig, FixCameraSet
        ucfg = _write_checkpoints(str(tmp_path / "model"))
        _write_envmaps(str(tmp_path / "envmap"))
        obj = tmp_path / "apple.obj"
        _write_obj(str(obj))

        # the two plugins the yaml keeps from the reference (prompt processor: cached text embeddings; background: unused
        # by the raytracing renderer) as minimal stand-ins
        @threestudio.register("solid-color-background")
        class _Bg(BaseModule):
            pass

        @threestudio.register("stable-diffusion-prompt-processor")
        class _PP(BaseObject):
            from dataclasses import dataclass as _dc

            @_dc
            class Config:
                prompt: str = ""

            def configure(self):
                g = torch.Generator().manual_seed(3)
                D = ucfg.cross_attention_dim
                vd, uvd, null = torch.randn(4, 77, D, generator=g), torch.randn(4, 77, D, generator=g), torch.randn(1, 77, D, generator=g)
                self.out = PromptProcessorOutput(vd[:1], uvd[:1], null, vd, uvd)

            def __call__(self):
                return self.out

        res = 64
        cfg = dict(YAML_SYSTEM)
        cfg["geometry"] = dict(cfg["geometry"], shape_init=f"mesh:{obj}")
        cfg["material"] = dict(cfg["material"], environment_texture=str(tmp_path / "envmap"))
        cfg["guidance"] = dict(cfg["guidance"], cache_dir=str(tmp_path / "model"), controlnet_path=str(tmp_path / "model" / "controlnet"))
        system = threestudio.find("dreammat-system")(cfg, resumed=False)
        for name in FIVE:
            assert threestudio.find(name).__module__ == "dreammat_b200.threestudio_plugin"
        assert type(system.geometry) is threestudio.find("dreammat-mesh") and type(system.renderer) is threestudio.find("raytracing-renderer")
        assert system.renderer.geometry is system.geometry and system.renderer.material is system.material
        assert "renderer.geometry.encoding.encoding.encoding.params" not in system.state_dict()      # sub-modules stay un-registered
        assert "geometry.encoding.encoding.encoding.params" in system.state_dict() and "renderer.bbox" in system.state_dict()
        system.on_fit_start()
        assert type(system.guidance) is threestudio.find("stable-diffusion-dreammat-guidance")
        system.impl.resize_to_vae = False         # 64^2 render straight into the (tiny) VAE keeps the test in seconds
        opt = system.configure_optimizers()["optimizer"]
        assert isinstance(opt, torch.optim.Adam) and opt.defaults["eps"] == 1e-15 and opt.defaults["betas"] == (0.9, 0.99)
        assert sum(p.numel() for g_ in opt.param_groups for p in g_["params"] if p.requires_grad) == 12599920 + 64 * 32 + 5 * 64
        cams = FixCameraSet(DataConfig(width=res, height=res), torch.Generator().manual_seed(0))
        gsel = torch.Generator().manual_seed(1)
        losses, p_hist = [], [system.geometry.impl.params.clone()]
        for step in range(3):
            vid, eid = cams.collate(gsel, 1)
            batch = {k: (x.cuda() if torch.is_tensor(x) else x) for k, x in cams.cameras(vid).items()}
            batch.update(view_id=vid, env_id=eid, condition_map=torch.rand(1, res, res, 22, device="cuda"), light_positions=None,
                         camera_positions=None)
            system.global_step = step
            system.on_train_batch_start(batch, step)                    # Updateable walk -> guidance.update_step(epoch, step)
            if step == 2:
                system.cfg.fused_step = False                           # the op-by-op autograd path gives gradients the same way
            out = system.training_step(batch, step)
            opt.zero_grad()
            out["loss"].backward()
            gp = system.geometry.encoding.encoding.encoding.params.grad
            assert gp is not None and torch.isfinite(gp).all() and float(gp.abs().sum()) > 0
            for p in (system.geometry.feature_network.layers[0].weight, system.geometry.feature_network.layers[2].weight):
                assert p.grad is not None and float(p.grad.abs().sum()) > 0
            opt.step()
            system.on_train_batch_end(out, batch, step)
            losses.append(float(out["loss"]))
            p_hist.append(system.geometry.impl.params.clone())
        assert all(np.isfinite(losses))
        for a, b in zip(p_hist[:-1], p_hist[1:]):
            assert not torch.equal(a, b)                                 # the optimizer moved the implementation's flat buffer
        assert system.guidance.impl.uncond_scale == pytest.approx(-1.0 + 0.5 * 2 / 2000)    # schedules follow the trainer's step
        # a checkpoint written by this system loads into a fresh geometry with strict=True (reference key names)
        sd = {k[len("geometry."):]: v for k, v in system.state_dict().items() if k.startswith("geometry.")}
        geo2 = threestudio.find("dreammat-mesh")(cfg["geometry"])
        geo2.load_state_dict(sd, strict=True)
        assert torch.equal(geo2.impl.params, system.geometry.impl.params)
    finally:
        sys.path.remove(STUB)
        for m in [k for k in sys.modules if k == "threestudio" or k.startswith("threestudio.") or k == "dreammat_b200.threestudio_plugin"]:
            del sys.modules[m]
