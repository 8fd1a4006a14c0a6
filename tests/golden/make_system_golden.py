"""Golden record of the system-level training step (a9 / boundary), made by EXECUTING the reference's own
`DreamMat.training_step` (systems/dreammat.py:57-179) and `BaseSystem.C` (systems/base.py:92-93) with recording stand-ins for
the renderer, the guidance, Lightning's `self.log` and the saver's `self.save_image_grid`.

What it pins: which keys the step hands to the guidance (`cond_normal`, `cond_depth` added to the batch, `rgb_as_latents=False`),
the loss assembly (`loss_* x C(lambda_*)` over the guidance outputs, then over the renderer outputs), the order and names of
everything logged, and the layout of the train-image grid (two rows: eight render outputs, eight channel groups of the
22-channel condition map); and, for `validation_step` / `test_step` / `on_test_epoch_end` (:181-300), the file names, grid
layouts, the four RGBA maps written per test view and the turntable gif call.

Run from the repo root where /root/reference exists:  python tests/golden/make_system_golden.py -> system_vectors.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden.make_golden import REF, Fake, base_ns, lift     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "system_vectors.pt")


def scenario(step, seed=5, H=6, W=5):
    """inputs shared by the generator and the test: renderer outputs, guidance outputs, batch"""
    g = torch.Generator().manual_seed(seed + step)
    out = {k: torch.rand(2, H, W, c, generator=g) for k, c in (("comp_rgb", 3), ("opacity", 1), ("comp_depth", 1), ("comp_normal", 3), ("albedo", 3),
                                                               ("metalness", 1), ("roughness", 1), ("specular_light", 3), ("diffuse_light", 3),
                                                               ("specular_color", 3), ("diffuse_color", 3))}
    out["loss_mat_reg"] = torch.rand((), generator=g)
    gout = {"loss_sds": torch.rand((), generator=g) * 10, "grad_norm": torch.rand((), generator=g)}
    for k in ("uncond_m_noise_norm", "text_m_noise_norm", "text_m_uncond_norm", "text_m_null_norm", "null_m_uncond_norm", "noise_norm",
              "uncond_norm", "text_norm"):
        gout[k] = torch.rand((), generator=g)
    batch = {"condition_map": torch.rand(2, H, W, 22, generator=g), "elevation": torch.tensor([1.0, 2.0]), "env_id": torch.tensor([0, 1])}
    return out, gout, batch


LOSS = {"lambda_sds": [0, 1.0, 0.5, 1000], "lambda_mat_reg": 10.0}      # one scheduled, one constant weight


def main():
    nc = base_ns(); nc["config_to_primitive"] = lambda v: list(v)
    lift("utils/misc.py", ["C"], nc)
    ns = base_ns()
    lift("systems/base.py", ["C"], ns, cls="BaseSystem")
    ns["C_method"] = ns["C"]            # BaseSystem.C; the free function C it calls (utils/misc.py) takes the name back
    ns["C"] = nc["C"]
    lift("systems/dreammat.py", ["training_step", "validation_step", "test_step", "on_test_epoch_end"], ns, cls="DreamMat")
    records = []
    for step, save_iter in ((0, 1), (400, 1000), (3000, 1000)):
        out, gout, batch = scenario(step)
        log, grids, seen = [], [], {}

        class Sys(Fake):
            def __call__(self, b):
                seen["renderer_batch_keys"] = sorted(b)
                return out
        me = Sys(cfg=Fake(loss=dict(LOSS), save_train_image=True, save_train_image_iter=save_iter, texture=True), true_global_step=step,
                 true_current_epoch=0, prompt_processor=lambda: "PROMPT_UTILS", log=lambda name, value: log.append((name, float(value))))

        def guidance(rgb, prompt_utils, **kw):
            seen["guidance"] = dict(rgb_is_comp_rgb=rgb is out["comp_rgb"], prompt_utils=prompt_utils, keys=sorted(kw),
                                    cond_normal_is_comp_normal=kw.get("cond_normal") is out["comp_normal"],
                                    cond_depth_is_comp_depth=kw.get("cond_depth") is out["comp_depth"], rgb_as_latents=kw.get("rgb_as_latents"))
            return gout
        me.guidance = guidance
        me.save_image_grid = lambda fn, imgs=None, name=None, step=None: grids.append(dict(
            filename=fn, name=name, step=step, rows=[[dict(type=c["type"], kwargs=c["kwargs"], img=c["img"].clone()) for c in row] for row in imgs]))
        me.C = lambda v, _me=me: ns["C_method"](_me, v)
        me.bind(ns, ["training_step"])
        ret = me.training_step(dict(batch), 0)
        records.append(dict(step=step, save_iter=save_iter, loss=float(ret["loss"]), ret_keys=sorted(ret), log=log, grids=grids, seen=seen))
    # validation / test hooks (:181-300): grids per view, the four RGBA maps the texture baker reads, the turntable gif
    evals = []
    for texture in (True, False):
        out, _, batch = scenario(77)
        batch["index"] = torch.tensor([7])
        calls = []
        me = Fake(cfg=Fake(texture=texture), true_global_step=1234)
        me.__class__ = type("S", (Fake,), {"__call__": lambda self, b: out})
        me.save_image_grid = lambda fn, imgs=None, name=None, step=None: calls.append(("grid", fn, name, step, [dict(type=c["type"], kwargs=c["kwargs"], img=c["img"].clone()) for c in imgs]))
        me.save_img = lambda img, fn: calls.append(("img", fn, img.clone()))
        me.save_gif = lambda path, fps=None: calls.append(("gif", path, fps))
        me.bind(ns, ["validation_step", "test_step", "on_test_epoch_end"])
        me.validation_step(batch)
        me.test_step(batch)
        me.on_test_epoch_end()
        evals.append(dict(texture=texture, calls=calls))
    torch.save({"loss_cfg": LOSS, "records": records, "evals": evals}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [(r["step"], round(r["loss"], 4), len(r["log"]), len(r["grids"])) for r in records])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    main()
