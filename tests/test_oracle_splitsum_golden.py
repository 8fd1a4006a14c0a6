"""The oracle's split-sum branch (a5) against vectors produced by executing the reference's own `DreamMatMaterial.forward`
(use_raytracing=False) and `shade_splitsum` (dreammat_material.py:679-711,747-762) on the real FG LUT -- see
tests/golden/make_splitsum_golden.py for what was executed and which two native calls were stood in for."""
import os

import torch

from oracle import render as O
from tests._fixtures import analytic_envlight, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 2e-6


def _load():
    return torch.load(os.path.join(HERE, "golden", "splitsum_vectors.pt")), torch.load(os.path.join(HERE, "golden", "splitsum_assets.pt"))["fg_lut"][0]


def test_splitsum_branch_matches_reference_execution():
    G, lut = _load()
    i, o = G["in"], G["out"]
    assert G["n_back_facing"] >= 5 and G["n_clamped_channels"] > 0          # the clamps are exercised
    f = i["features"].clone().requires_grad_(True)
    fj = i["features_jitter"].clone().requires_grad_(True)
    albedo, metallic, rough, reg = O.material_params(f, fj, use_raytracing=False)
    out = O.shade_splitsum_with(i["normals"], i["viewdirs"], lambda ndv, r: O.fg_lookup(lut, ndv, r), analytic_envlight,
                                analytic_envlight, metallic, rough, albedo)
    ((out["color"] * i["cotangent"]).sum() + i["reg_weight"] * reg).backward()
    errs = {k: rel_err(out[k].detach(), o[k]) for k in ("color", "albedo", "roughness", "metalness", "specular_lights", "diffuse_lights",
                                                        "specular_colors", "diffuse_colors")}
    errs["mat_reg"] = abs(float(reg) - float(o["mat_reg"])) / abs(float(o["mat_reg"]))
    errs["d_features"] = rel_err(f.grad, o["d_features"])
    errs["d_features_jitter"] = rel_err(fj.grad, o["d_features_jitter"])
    # the gradient w.r.t. the features passes through d LUT / d uv = 256 x (difference of neighbouring texels): grid_sample's
    # backward and autograd through fg_lookup round that cancellation differently (2e-6 measured)
    g_tol = {"d_features": 1e-5}
    assert all(v < g_tol.get(k, TOL) for k, v in errs.items()), errs
    # the split-sum branch uses the LINEAR roughness range (min_roughness .. max_roughness), unlike the ray-traced branch
    assert float(out["roughness"].min()) >= 0.1 and float(out["roughness"].max()) <= 0.95 and float(out["roughness"].max()) > 0.9


def test_fg_lookup_matches_an_independent_bilinear_clamp():
    """fg values the reference run fetched (through grid_sample, texel centres at (i + 0.5) / 256, clamped) are recovered from
    the stored outputs: specular_albedo = F0 * fg.x + fg.y is linear in F0, so two pixels are not needed -- compare the LUT
    fetch directly on the run's own coordinates instead, including the back-facing (clamped to 0) ones."""
    G, lut = _load()
    i = G["in"]
    ndv = (i["normals"] * i["viewdirs"]).sum(-1)
    rough = G["out"]["roughness"][:, 0]
    uv = torch.stack([ndv, rough], -1).clamp(0, 1).view(1, -1, 1, 2)
    ref = torch.nn.functional.grid_sample(lut[None].permute(0, 3, 1, 2), uv * 2 - 1, mode="bilinear", padding_mode="border",
                                          align_corners=False)[0, :, :, 0].t()
    assert rel_err(O.fg_lookup(lut, ndv, rough), ref) < TOL
    edge = torch.tensor([0.0, 1.0, 0.0, 1.0]), torch.tensor([0.0, 0.0, 1.0, 1.0])                    # the four corners
    uv = torch.stack(edge, -1).view(1, -1, 1, 2)
    ref = torch.nn.functional.grid_sample(lut[None].permute(0, 3, 1, 2), uv * 2 - 1, mode="bilinear", padding_mode="border",
                                          align_corners=False)[0, :, :, 0].t()
    assert torch.allclose(O.fg_lookup(lut, *edge), ref, atol=1e-7)
