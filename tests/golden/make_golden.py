"""Generate golden input/output vectors by EXECUTING the reference's own function bodies (CPU, torch).

The reference package cannot be imported as a whole here (pytorch_lightning, omegaconf, diffusers, nvdiffrast,
tiny-cuda-nn, envlight, jaxtyping ... are absent), but most of the arithmetic of the hot path lives in plain-torch
functions and methods.  This script lifts those function definitions out of the files under /root/reference by AST
(annotations and decorators stripped, nothing else touched), executes them on seeded inputs with a minimal fake `self`
where they are methods, and stores inputs + outputs in tests/golden/reference_vectors.pt.  The CPU oracle (oracle/) is
then pinned against these vectors by tests/test_oracle_golden.py -- on any machine, without /root/reference.

Run from the repo root (only where /root/reference exists):  python tests/golden/make_golden.py

What is covered (reference file:line -> golden key):
  utils/ops.py:179-292, data/uncond.py:723-821            -> "collate"      (a1: cameras, rays, mvp, view/env draws)
  models/geometry/base.py:20-32, utils/ops.py:26-37       -> "contract"     (a3: contract_to_unisphere)
  models/renderers/raytracing_renderer.py:161-173,306-343 -> "jitter", "controlnet_maps" (a2/a3)
  models/materials/dreammat_material.py:89-123,490-677,713-797 -> "material" (a4 forward + autograd backward, export)
  models/guidance/dreammat_guidance.py:440-497,584-602    -> "guidance"     (a8/a9: CSD combination, loss_sds, its gradient)
  utils/misc.py:65-86                                     -> "C"            (schedules)
  models/mesh.py:135-161                                  -> "vertex_normals"
Not coverable by execution (their arithmetic is inside absent native packages): tiny-cuda-nn hash grid, nvdiffrast
rasterize / antialias / texture, envlight cubemaps, diffusers UNet / ControlNet / VAE, the `_raytracing` BVH.
"""
import ast
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference/threestudio_dreammat/threestudio"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.pt")


class _Strip(ast.NodeTransformer):
    """Remove annotations / decorators so the bodies run without jaxtyping & co."""

    def visit_FunctionDef(self, node):
        self.generic_visit(node)
        node.decorator_list = []
        node.returns = None
        for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs:
            a.annotation = None
        if node.args.vararg:
            node.args.vararg.annotation = None
        if node.args.kwarg:
            node.args.kwarg.annotation = None
        return node

    def visit_AnnAssign(self, node):
        self.generic_visit(node)
        if node.value is None:
            return None
        return ast.copy_location(ast.Assign(targets=[node.target], value=node.value), node)


def base_ns():
    return {"torch": torch, "np": np, "F": F, "nn": nn, "math": math, "Tensor": torch.Tensor, "os": os}


def lift(path, names, ns, cls=None):
    """exec the named function definitions of a reference file (top-level, or methods of class `cls`) into ns."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    if cls is not None:
        tree = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls)
    found = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            mod = ast.Module(body=[_Strip().visit(node)], type_ignores=[])
            ast.fix_missing_locations(mod)
            exec(compile(mod, os.path.join(REF, path), "exec"), ns)
            found.add(node.name)
    missing = set(names) - found
    assert not missing, (path, missing)
    return ns


def lift_block(path, first_marker, last_marker, ns, fn_name, args, ret):
    """exec an inline block of a reference method (from the line containing first_marker to the one containing
    last_marker) as the body of a function fn_name(args) returning `ret`."""
    lines = open(os.path.join(REF, path)).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if first_marker in l)
    i1 = next(i for i, l in enumerate(lines) if last_marker in l and i >= i0)
    body = lines[i0:i1 + 1]
    ind = min(len(l) - len(l.lstrip()) for l in body if l.strip())
    src = f"def {fn_name}({', '.join(args)}):\n" + "\n".join("    " + l[ind:] for l in body) + f"\n    return {ret}\n"
    exec(compile(src, os.path.join(REF, path) + f":{i0 + 1}-{i1 + 1}", "exec"), ns)
    return (i0 + 1, i1 + 1)


class Fake:
    """bare object used as `self`; reference methods are bound to it"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def bind(self, ns, names):
        for n in names:
            setattr(self, n, types.MethodType(ns[n], self))
        return self


def sphere_tracer(center, radius):
    """analytic occluder standing in for the BVH (`ray_trace_fun(o, d) -> inters, normals, depth, hit_mask`,
    raytracing_renderer.py:318-323 semantics: depth 10 on a miss)"""
    c = torch.tensor(center, dtype=torch.float32)

    def fn(o, d):
        oc = o - c
        b = (oc * d).sum(-1)
        cc = (oc * oc).sum(-1) - radius * radius
        disc = b * b - cc
        t = -b - torch.sqrt(disc.clamp_min(0))
        hit = (disc > 0) & (t > 0)
        depth = torch.where(hit, t, torch.full_like(t, 10.0))
        inters = o + d * depth[:, None]
        nrm = F.normalize(inters - c, dim=-1)
        return inters, nrm, depth[:, None], hit
    return fn


def main():
    G = {}
    torch.manual_seed(0)

    # ------------------------------------------------------------------ ops + collate (a1)
    ns = base_ns()
    lift("utils/ops.py", ["dot", "reflect", "scale_tensor", "get_activation", "get_ray_directions", "get_rays",
                          "get_projection_matrix", "get_mvp_matrix"], ns)
    lift("data/uncond.py", ["collate"], ns, cls="FixCameraIterableDataset")
    H = W = 12
    NV, NE, B = 6, 3, 4
    g = torch.Generator().manual_seed(11)
    ds = Fake(batch_size=B, cfg=Fake(fix_view_num=NV, fix_env_num=NE), height=H, width=W,
              elevation_degs=torch.rand(NV, generator=g) * 65 - 20, azimuth_degs=torch.rand(NV, generator=g) * 360 - 180,
              fix_camera_distances=torch.rand(NV, generator=g) + 3, camera_perturbs=torch.zeros(NV, 3),
              center_perturbs=torch.zeros(NV, 3), up_perturbs=torch.zeros(NV, 3), fovy_degs=torch.rand(NV, generator=g) * 20 + 25,
              directions_unit_focal=ns["get_ray_directions"](H=H, W=W, focal=1.0),
              depths=torch.rand(NV, H, W, 1, generator=g), normals=torch.rand(NV, H, W, 3, generator=g),
              lightmaps=torch.rand(NV, NE, H, W, 18, generator=g)).bind(ns, ["collate"])
    torch.manual_seed(123)
    out = ds.collate(None)
    G["collate"] = {"in": {k: getattr(ds, k) for k in ("elevation_degs", "azimuth_degs", "fix_camera_distances", "fovy_degs", "depths",
                                                      "normals", "lightmaps")} | {"H": H, "W": W, "seed": 123, "B": B},
                    "out": {k: v for k, v in out.items() if torch.is_tensor(v)}}

    # ------------------------------------------------------------------ contract_to_unisphere (a3)
    lift("models/geometry/base.py", ["contract_to_unisphere"], ns)
    x = torch.rand(50, 3, generator=g) * 2.4 - 1.2
    bbox = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    G["contract"] = {"x": x, "bbox": bbox, "out": ns["contract_to_unisphere"](x.clone(), bbox),
                     "out_unbounded": ns["contract_to_unisphere"](x.clone(), bbox, True)}

    # ------------------------------------------------------------------ renderer helpers + jitter block (a2/a3)
    nr = base_ns()
    lift("models/renderers/raytracing_renderer.py", ["xfm_vectors"], nr)
    lift("models/renderers/raytracing_renderer.py", ["get_orthogonal_directions", "compute_controlnet_normals", "compute_controlnet_depth"], nr,
         cls="RaytraceRender")
    span = lift_block("models/renderers/raytracing_renderer.py", "x = self.get_orthogonal_directions(gb_normal[selector])",
                      "positions_jitter = gb_pos[selector] + change", nr, "jitter_block", ["self", "gb_pos", "gb_normal", "selector", "positions"],
                      "positions_jitter")
    ren = Fake(device="cpu", change_type="gaussian", change_eps=0.05).bind(
        nr, ["get_orthogonal_directions", "compute_controlnet_normals", "compute_controlnet_depth", "jitter_block"])
    Bn, P = 1, 40
    gb_pos = torch.rand(Bn, P, 3, generator=g) * 1.6 - 0.8
    gb_normal = F.normalize(torch.randn(Bn, P, 3, generator=g), dim=-1)
    gb_normal[0, 0] = torch.tensor([0.0, 0.0, 1.0]); gb_normal[0, 1] = torch.tensor([1.0, 0.0, 0.0])   # both branches of the mask
    selector = torch.rand(Bn, P, generator=g) > 0.3
    torch.manual_seed(321)
    pj = ren.jitter_block(gb_pos, gb_normal, selector, gb_pos[selector])
    G["jitter"] = {"gb_pos": gb_pos, "gb_normal": gb_normal, "selector": selector, "seed": 321, "positions_jitter": pj,
                   "ortho": ren.get_orthogonal_directions(gb_normal[selector]), "lines": span}
    nrm_hw = F.normalize(torch.randn(30, 3, generator=g), dim=-1)      # [selected pixels, 3] (raytracing_renderer.py:146)
    w2c = torch.eye(4)[None].clone(); w2c[0, :3, :3] = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
    depth = torch.rand(1, 6, 5, 1, generator=g) * 3 + 1
    hit = torch.rand(1, 6, 5, 1, generator=g) > 0.4
    G["controlnet_maps"] = {"normals": nrm_hw, "w2c": w2c, "normal_out": ren.compute_controlnet_normals(nrm_hw.clone(), w2c, 1),
                            "depth": depth, "hit": hit, "depth_out": ren.compute_controlnet_depth(depth.clone(), hit)}

    # ------------------------------------------------------------------ material (a4): forward + backward + export
    nm = base_ns()
    nm["get_activation"] = ns["get_activation"]
    lift("models/materials/dreammat_material.py", ["saturate_dot", "sample_sphere", "material_smoothness_grad"], nm)
    lift("models/materials/dreammat_material.py", [
        "get_envirmentlight_blender", "get_lights", "fresnel_schlick",
        "fresnel_schlick_directions", "geometry_schlick_ggx", "geometry_schlick", "get_orthogonal_directions", "sample_diffuse_directions",
        "sample_specular_directions", "distribution_ggx", "geometry", "shade_raytracing", "forward", "export", "set_raytracer"], nm, cls="DreamMatMaterial")
    ND, NS = 24, 16
    cfg = Fake(use_raytracing=True, material_activation="sigmoid", min_metallic=0.0, max_metallic=0.9, min_roughness_squre=0.01,
               max_roughness_squre=0.9, min_roughness=0.1, max_roughness=0.95, random_azimuth=True, geometry_type="schlick", use_bump=False,
               diffuse_sample_num=ND, specular_sample_num=NS)
    env = torch.rand(16, 32, 3, generator=g) * 2.0
    mat = Fake(cfg=cfg, light=[env])
    for name, n in (("diffuse_direction_samples", ND), ("specular_direction_samples", NS)):   # dreammat_material.py:388-398 (CPU)
        az, el = nm["sample_sphere"](n, 0)
        az, el = az * 0.5 / np.pi, 1 - 2 * el / np.pi
        setattr(mat, name, torch.from_numpy(np.stack([az, el], -1).astype(np.float32)))
    mat.bind(nm, ["get_envirmentlight_blender", "get_lights", "fresnel_schlick", "fresnel_schlick_directions", "geometry_schlick_ggx",
                  "geometry_schlick", "get_orthogonal_directions", "sample_diffuse_directions", "sample_specular_directions",
                  "distribution_ggx", "geometry", "shade_raytracing", "forward", "export", "set_raytracer"])
    occ = {"center": [0.35, 0.1, 0.9], "radius": 0.45}
    mat.set_raytracer(sphere_tracer(occ["center"], occ["radius"]))
    PN = 37
    nrm = F.normalize(torch.randn(PN, 3, generator=g), dim=-1)
    pts = nrm * 0.8 * 0.5
    vd = F.normalize(nrm + 0.7 * torch.randn(PN, 3, generator=g), dim=-1)
    feat = torch.randn(PN, 5, generator=g).requires_grad_(True)
    featj = (feat.detach() + 0.3 * torch.randn(PN, 5, generator=g)).requires_grad_(True)
    SEED = 777
    torch.manual_seed(SEED)
    outputs, mat_reg = mat.forward(pts, feat, featj, vd, nrm, 0)
    cot = torch.randn(PN, 3, generator=g)
    ((outputs["color"] * cot).sum() + 1.7 * mat_reg).backward()
    torch.manual_seed(SEED)          # the two draws shade_raytracing made, in order (:566, :589)
    rd = torch.rand((PN, 1, 1)); rs = torch.rand((PN, 1, 1))
    G["material"] = {"in": {"pts": pts, "normals": nrm, "viewdirs": vd, "features": feat.detach(), "features_jitter": featj.detach(),
                            "env": env, "occluder": occ, "rand_d": rd, "rand_s": rs, "cotangent": cot, "reg_weight": 1.7,
                            "n_diffuse": ND, "n_specular": NS, "tab_d": mat.diffuse_direction_samples, "tab_s": mat.specular_direction_samples},
                     "out": {k: v.detach() for k, v in outputs.items()} | {"mat_reg": mat_reg.detach(), "d_features": feat.grad.clone(),
                                                                          "d_features_jitter": featj.grad.clone()},
                     "export": {k: v.detach() for k, v in mat.export(feat.detach()).items()}}

    # ------------------------------------------------------------------ guidance: CSD combination + loss (a8/a9)
    ng = base_ns()
    lift("models/guidance/dreammat_guidance.py", ["compute_grad_sds", "__call__"], ng, cls="StableDiffusionLightGuidance")
    Bg = 3
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2       # SD scaled-linear schedule (published)
    alphas = torch.cumprod(1.0 - betas, 0)
    preds = [torch.randn(Bg, 4, 8, 8, generator=g) for _ in range(3)]
    lat0 = torch.randn(Bg, 4, 8, 8, generator=g)

    class Sched:
        def add_noise(self, x, n, t):
            a = alphas[t].view(-1, 1, 1, 1)
            return a.sqrt() * x + (1 - a).sqrt() * n
    gd = Fake(min_step=20, max_step=980, device="cpu", scheduler=Sched(), alphas=alphas, cond_scale=1.05, uncond_scale=-0.7,
              null_scale=-0.2, noise_scale=0.0, perpneg_scale=0.0, use_controlnet=False,
              cfg=Fake(grad_clip_val=None, grad_normalize=False, control_types=[], condition_scales=[]))
    gd.compute_without_perpneg = lambda *a, **k: tuple(preds)
    lat = lat0.clone().requires_grad_(True)
    gd.get_latents = lambda rgb_BCHW, rgb_as_latents=False: lat
    gd.bind(ng, ["compute_grad_sds", "__call__"])
    torch.manual_seed(99)
    gout = gd.__call__(torch.zeros(Bg, 8, 8, 3), Fake(use_perp_neg=False), torch.zeros(Bg), torch.zeros(Bg), torch.ones(Bg), torch.zeros(Bg))
    gout["loss_sds"].backward()
    torch.manual_seed(99)
    t = torch.randint(20, 981, [Bg], dtype=torch.long); noise = torch.randn_like(lat0)
    G["guidance"] = {"in": {"latents": lat0, "eps_text": preds[0], "eps_uncond": preds[1], "eps_null": preds[2], "t": t, "noise": noise,
                            "alphas": alphas, "scales": (1.05, -0.7, -0.2, 0.0)},
                     "out": {k: v.detach() for k, v in gout.items()} | {"d_latents": lat.grad.clone()}}

    # ------------------------------------------------------------------ schedules
    nc = base_ns(); nc["config_to_primitive"] = lambda v: list(v)
    lift("utils/misc.py", ["C"], nc)
    cases = [([0, -1.0, -0.5, 2000], 0, 0), ([0, -1.0, -0.5, 2000], 0, 1000), ([0, -1.0, -0.5, 2000], 0, 5000), ([500, 0.2, 0.02, 501], 0, 500),
             ([500, 0.2, 0.02, 501], 0, 501), (1.05, 0, 10), ([0.1, 0.9, 300], 0, 150), ([0, 1.0, 0.0, 2.0], 1, 77)]
    G["C"] = [(v, e, s, float(nc["C"](v, e, s))) for (v, e, s) in cases]

    # ------------------------------------------------------------------ vertex normals
    nv = base_ns(); nv["dot"] = ns["dot"]
    lift("models/mesh.py", ["_compute_vertex_normal"], nv, cls="Mesh")
    v = torch.randn(30, 3, generator=g)
    f = torch.randint(0, 30, (50, 3), generator=g)
    mesh = Fake(v_pos=v, t_pos_idx=f).bind(nv, ["_compute_vertex_normal"])
    G["vertex_normals"] = {"v": v, "f": f, "out": mesh._compute_vertex_normal()}

    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; keys:", sorted(G))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: golden vectors can only be regenerated where /root/reference exists")
    main()
