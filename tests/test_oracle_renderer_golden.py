"""The oracle's renderer composition (a2 / a3 / a6) against vectors produced by executing the reference's own
`RaytraceRender.forward` (raytracing_renderer.py:110-222) with its rasteriser wrapper and its material -- see
tests/golden/make_renderer_golden.py for what was executed and what stood in for the absent native packages."""
import os

import torch

from oracle import render as O

HERE = os.path.dirname(os.path.abspath(__file__))


def close(a, b, tol=1e-6):
    a, b = a.detach().double(), b.detach().double()
    return a.shape == b.shape and float((a - b).abs().max()) <= tol * (1.0 + float(b.abs().max()))


def _sphere_hit(occ):
    c = torch.tensor(occ["center"], dtype=torch.float32)
    r = occ["radius"]

    def fn(o, d):
        oc = o - c
        b = (oc * d).sum(-1)
        disc = b * b - ((oc * oc).sum(-1) - r * r)
        t = -b - torch.sqrt(disc.clamp_min(0))
        return (disc > 0) & (t > 0)
    return fn


def test_renderer_forward_matches_reference_execution():
    G = torch.load(os.path.join(HERE, "golden", "renderer_vectors.pt"))
    gi, go, st = G["in"], G["out"], G["standins"]
    v, f, vn, mvp = gi["v"], gi["f"], gi["vn"], gi["mvp_mtx"]
    # the clip-space vertices the reference handed to dr.rasterize (NVDiffRasterizerContext.vertex_transform) are the oracle's
    clip = torch.cat([v, torch.ones_like(v[:, :1])], -1) @ mvp[0].t()
    assert close(clip[None], st["clip_seen_by_rasterize"])
    tracer = O.RayTracer(v.numpy(), f.numpy())
    gb = O.gbuffer(tracer, v, f, vn, gi["rays_o"], gi["rays_d"], mvp, gi["w2c"])
    assert torch.equal(gb["rast"], st["rast"])                    # the stand-in rasteriser IS this function (deterministic C tracer)
    pairs = O.antialias_pairs(gb["rast"][0], v, f, mvp[0])
    assert len(pairs) == st["n_aa_pairs"] > 0
    meta, n_entries = O.hashgrid_meta(n_levels=gi["hash_levels"], log2_T=gi["hash_log2_T"])
    assert n_entries * 2 == gi["grid"].numel()

    def geometry(p):
        return O.geometry_forward(p, gi["grid"], gi["W1"], gi["W2"], meta)

    def material(pts, fe, fj, vd, nrm):
        al, me, ro, reg = O.material_params(fe, fj)
        out = O.shade_raytracing(pts, nrm, vd, gi["env"], me, ro, al, gi["rand_d"], gi["rand_s"], _sphere_hit(gi["occluder"]),
                                 n_diffuse=gi["n_diffuse"], n_specular=gi["n_specular"])
        assert 0.02 < float(out["_hit"].float().mean()) < 0.98    # the occluder matters
        return out, reg
    out = O.render_forward(gb, pairs, geometry, material, gi["rand_ang"], gi["normal_eps"])
    assert set(out) == set(go) and len(go) == 12
    assert int(gb["selector"].sum()) == G["pn"]
    for k in sorted(go):
        tol = 5e-6 if k in ("comp_rgb", "specular_light", "diffuse_light", "specular_color") else 2e-6
        assert close(out[k], go[k], tol), (k, tuple(out[k].shape), tuple(go[k].shape), float((out[k].double() - go[k].double()).abs().max()))
    # properties of the reference's composition the product relies on
    H = gi["res"]
    m = gb["mask"][0, ..., 0]
    assert float(go["comp_depth"][0][~m].abs().max()) == 0.0 and float(go["comp_depth"][0][m].min()) >= 0.3 - 1e-6
    assert abs(float(go["comp_depth"][0][m].max()) - 1.0) < 1e-4                     # nearest covered pixel -> 1, farthest -> 0.3
    untouched = torch.ones(H * H, dtype=torch.bool)
    untouched[m.reshape(-1)] = False
    for (dst, src, _) in pairs:
        untouched[dst] = False
    assert torch.equal(go["comp_rgb"].reshape(-1, 3)[untouched], torch.ones(int(untouched.sum()), 3))   # canvas of ones
    assert torch.allclose(go["comp_normal"].reshape(-1, 3)[untouched], torch.tensor([0.5, 0.5, 1.0]).expand(int(untouched.sum()), 3))
