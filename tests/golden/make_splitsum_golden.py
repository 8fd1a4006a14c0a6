"""Golden vectors for the split-sum branch (a5), made by EXECUTING the reference's own code.

`DreamMatMaterial.forward` (the `use_raytracing=False` branch, dreammat_material.py:747-762) and `shade_splitsum`
(:679-711) are lifted from /root/reference by AST (tests/golden/make_golden.py's machinery) and run on seeded inputs.
Their two native dependencies are absent here and are replaced by stand-ins, stated once:

  dr.texture(FG_LUT, uv, 'linear', 'clamp')  -> torch.nn.functional.grid_sample(align_corners=False, padding_mode='border'):
                                                texel centres at (i + 0.5) / N, bilinear, clamped -- nvdiffrast's published rule
  self.envlight[env_id](l[, roughness])      -> tests/_fixtures.analytic_envlight (a smooth closed form)

so what the vectors pin is everything the REFERENCE wrote for this branch: the material ranges of the split-sum branch
(min_roughness / max_roughness, not the squared ones), mat_reg, n.v, the reflection vector, the (n.v, roughness) order of the
LUT coordinates and their clamp, F0, the specular albedo, the colour clamp, the eight outputs and the autograd gradients --
on the REAL bsdf_256_256.bin (tests/golden/splitsum_assets.pt).  The oracle's `fg_lookup` is thereby also checked against an
independent bilinear implementation.

Run from the repo root where /root/reference exists:  python tests/golden/make_splitsum_golden.py -> splitsum_vectors.pt
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests._fixtures import analytic_envlight                      # noqa: E402
from tests.golden.make_golden import REF, Fake, base_ns, lift      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitsum_vectors.pt")
ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitsum_assets.pt")


class _Dr:
    """the one nvdiffrast call of the branch"""

    @staticmethod
    def texture(tex, uv, filter_mode="linear", boundary_mode="clamp"):
        assert filter_mode == "linear" and boundary_mode == "clamp" and tex.dim() == 4 and uv.shape[-1] == 2
        out = F.grid_sample(tex.permute(0, 3, 1, 2), uv * 2.0 - 1.0, mode="bilinear", padding_mode="border", align_corners=False)
        return out.permute(0, 2, 3, 1)


def main():
    ns = base_ns()
    lift("utils/ops.py", ["get_activation"], ns)
    nm = base_ns()
    nm["get_activation"] = ns["get_activation"]
    nm["dr"] = _Dr
    lift("models/materials/dreammat_material.py", ["material_smoothness_grad"], nm)
    lift("models/materials/dreammat_material.py", ["shade_splitsum", "forward"], nm, cls="DreamMatMaterial")
    lut = torch.load(ASSETS)["fg_lut"]                               # [1,256,256,2], read by the reference's own statement
    cfg = Fake(use_raytracing=False, material_activation="sigmoid", min_metallic=0.0, max_metallic=0.9, min_roughness_squre=0.01,
               max_roughness_squre=0.9, min_roughness=0.1, max_roughness=0.95)
    seen = {}

    def env(l, roughness=None):
        out = analytic_envlight(l, roughness)
        seen["specular_light" if roughness is not None else "diffuse_light"] = out.detach().clone()
        return out
    mat = Fake(cfg=cfg, FG_LUT=lut, envlight=[env]).bind(nm, ["shade_splitsum", "forward"])
    g = torch.Generator().manual_seed(23)
    PN = 53
    nrm = F.normalize(torch.randn(PN, 3, generator=g), dim=-1)
    vd = F.normalize(nrm + 0.9 * torch.randn(PN, 3, generator=g), dim=-1)
    vd[:5] = -vd[:5]                                                 # back-facing views: n.v < 0 -> the LUT coordinate clamps
    feat = (2.5 * torch.randn(PN, 5, generator=g)).requires_grad_(True)      # wide: roughness / metallic reach their range ends
    featj = (feat.detach() + 0.3 * torch.randn(PN, 5, generator=g)).requires_grad_(True)
    outputs, mat_reg = mat.forward(nrm * 0.4, feat, featj, vd, nrm, 0)
    cot = torch.randn(PN, 3, generator=g)
    ((outputs["color"] * cot).sum() + 2.3 * mat_reg).backward()
    G = {"in": {"normals": nrm, "viewdirs": vd, "features": feat.detach(), "features_jitter": featj.detach(), "cotangent": cot,
                "reg_weight": 2.3},
         "out": {k: v.detach() for k, v in outputs.items()} | {"mat_reg": mat_reg.detach(), "d_features": feat.grad.clone(),
                                                              "d_features_jitter": featj.grad.clone(), **seen},
         "n_back_facing": int(((nrm * vd).sum(-1) < 0).sum()), "n_clamped_channels": int(((outputs["color"] <= 0) | (outputs["color"] >= 1)).sum())}
    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; back-facing", G["n_back_facing"], "clamped colour channels", G["n_clamped_channels"])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    main()
