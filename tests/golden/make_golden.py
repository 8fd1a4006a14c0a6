"""Generate golden input/output vectors by EXECUTING the reference's own function bodies (CPU, torch).

The reference package cannot be imported as a whole here (pytorch_lightning, omegaconf, diffusers, nvdiffrast,
tiny-cuda-nn, envlight, jaxtyping ... are absent), but most of the arithmetic of the hot path lives in plain-torch
functions and methods.  This script lifts those function definitions out of the files under /root/reference by AST
(annotations and decorators stripped, nothing else touched), executes them on seeded inputs with a minimal fake `self`
where they are methods, and stores inputs + outputs in tests/golden/reference_vectors.pt.  The CPU oracle (oracle/) is
then pinned against these vectors by tests/test_oracle_golden.py -- on any machine, without /root/reference.

Run from the repo root (only where /root/reference exists):  python tests/golden/make_golden.py

What is covered (reference file:line -> golden key):
  data/uncond.py:584-645,692-698                          -> "fixed_views"  (a1: the 128 fixed cameras, draw order)
  models/materials/dreammat_material.py:65-68             -> "envmap"       (a4 input: EXR decoding of the shipped map1.exr)
  data/uncond.py:532-557 (loadrgb / loaddepth)            -> "maps"         (a1 / N1: condition-map PNG decoding)
  utils/ops.py:179-292, data/uncond.py:723-821            -> "collate"      (a1: cameras, rays, mvp, view/env draws)
  models/geometry/base.py:20-32, utils/ops.py:26-37       -> "contract"     (a3: contract_to_unisphere)
  models/renderers/raytracing_renderer.py:161-173,306-343 -> "jitter", "controlnet_maps" (a2/a3)
  models/materials/dreammat_material.py:89-123,490-677,713-797 -> "material" (a4 forward + autograd backward, export)
  models/guidance/dreammat_guidance.py:440-497,584-602    -> "guidance"     (a8/a9: CSD combination, loss_sds, its gradient)
  models/prompt_processors/base.py:52-85,187-189,243-309  -> "prompt"       (a8: view-dependent embedding selection, [text|uncond|null])
  models/guidance/dreammat_guidance.py:388-438            -> "branches"     (a8: CFG batch layout of the UNet call)
  models/networks.py:150-187                              -> "mlp"          (a3: VanillaMLP structure, keys, forward)
  utils/misc.py:65-86                                     -> "C"            (schedules)
  models/mesh.py:135-161                                  -> "vertex_normals"
  models/geometry/dreammat_mesh.py:163-197 (inline block) -> "mesh_normalize"
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


class _Sub:
    """stands in for jaxtyping's Float[Tensor, "..."] in dataclass field annotations"""

    def __class_getitem__(cls, item):
        return cls


def lift_class(path, cls, ns, keep_fields=False):
    """exec a whole class definition of a reference file into ns (function annotations stripped; keep_fields leaves
    the class-level annotated fields in place, for @dataclass definitions)."""
    src = open(os.path.join(REF, path)).read()
    node = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.ClassDef) and n.name == cls)
    if keep_fields:
        import dataclasses
        import typing
        ns.update({k: getattr(typing, k) for k in ("Callable", "List", "Dict", "Tuple", "Optional", "Any")})
        ns.update(dataclass=dataclasses.dataclass, field=dataclasses.field, Float=_Sub, Bool=_Sub, Int=_Sub)
        node.body = [(_Strip().visit(b) if isinstance(b, ast.FunctionDef) else b) for b in node.body]
        mod = ast.Module(body=[node], type_ignores=[])
    else:
        mod = ast.Module(body=[_Strip().visit(node)], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, os.path.join(REF, path), "exec"), ns)
    return ns[cls]


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

    # ------------------------------------------------------------------ the fixed view set drawn at dataset construction (a1)
    names = ["set_fix_elevs", "set_fix_azims", "set_fix_camera_distance", "set_fix_camera_perturb", "set_fix_center_perturb",
             "set_fix_up_perturb", "set_fix_fovy"]
    lift("data/uncond.py", names, ns, cls="FixCameraIterableDataset")
    fv = Fake(cfg=Fake(fix_view_num=128, camera_perturb=0.0, center_perturb=0.0, up_perturb=0.0), elevation_range=(-20, 45),
              azimuth_range=(-180, 180), camera_distance_range=(3, 4), fovy_range=(25, 45)).bind(ns, names)   # configs/dreammat.yaml:11-17
    torch.manual_seed(2024)
    for n_ in names:          # the call order of __init__ (uncond.py:692-698)
        getattr(fv, n_)()
    G["fixed_views"] = {"seed": 2024, "elevation_degs": fv.elevation_degs, "azimuth_degs": fv.azimuth_degs,
                        "camera_distances": fv.fix_camera_distances, "fovy_degs": fv.fovy_degs}

    # ------------------------------------------------------------------ condition-map files (a1 / N1): loadrgb, loaddepth
    import tempfile
    import cv2
    nl = base_ns(); nl["cv2"] = cv2
    lift("data/uncond.py", ["loadrgb", "loaddepth"], nl, cls="FixCameraIterableDataset")      # nested in render_fixview_imgs
    rng = np.random.RandomState(3)
    src_hw, dst_hw = 48, 32                                              # files are rendered larger than the training size
    yy, xx = np.mgrid[0:src_hw, 0:src_hw]
    disk = ((yy - 24) ** 2 + (xx - 22) ** 2) < 17 ** 2
    depth_mm = np.where(disk, 2500 + 40 * yy + 15 * xx + rng.randint(0, 30, (src_hw, src_hw)), 0).astype(np.uint16)
    rgb_files = {"normal": rng.randint(0, 256, (src_hw, src_hw, 3)).astype(np.uint8)}
    for tag in ("m0.0r0.0", "m0.0r0.5", "m0.0r1.0", "m1.0r0.0", "m1.0r0.5", "m1.0r1.0"):
        rgb_files[tag] = rng.randint(0, 256, (src_hw, src_hw, 3)).astype(np.uint8)
    with tempfile.TemporaryDirectory() as td:
        cv2.imwrite(os.path.join(td, "d.png"), depth_mm)
        d_out = nl["loaddepth"](os.path.join(td, "d.png"), (dst_hw, dst_hw))
        rgb_out = {}
        for k, arr in rgb_files.items():
            cv2.imwrite(os.path.join(td, "c.png"), arr)                  # cv2 writes the array as BGR
            rgb_out[k] = torch.from_numpy(nl["loadrgb"](os.path.join(td, "c.png"), (dst_hw, dst_hw)))
    G["maps"] = {"depth_png_u16": torch.from_numpy(depth_mm.astype(np.int32)), "rgb_png_u8": {k: torch.from_numpy(v) for k, v in rgb_files.items()},
                 "size": dst_hw, "depth": torch.from_numpy(d_out).float(), "rgb": rgb_out}

    # ------------------------------------------------------------------ env map decoding (a4 input): load_hdr_image on the shipped map1.exr
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    nh = base_ns(); nh["cv2"] = cv2
    lift("models/materials/dreammat_material.py", ["load_hdr_image"], nh)
    exr = os.path.join(os.path.dirname(REF), "load/lights/envmap/map1/map1.exr")
    img = torch.tensor(nh["load_hdr_image"](exr), dtype=torch.float32)            # configure(): torch.tensor(load_hdr_image(pathexr), float32)
    G["envmap"] = {"relpath": "load/lights/envmap/map1/map1.exr", "shape": tuple(img.shape), "sum": float(img.double().sum()),
                   "min": float(img.min()), "max": float(img.max()), "crop": img[1000:1016, 2000:2016].clone(),
                   "rows": img[::256, ::512].clone()}

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

    # ------------------------------------------------------------------ view-dependent prompt selection + CFG branch layout (a8)
    npp = base_ns()
    lift("models/prompt_processors/base.py", ["shift_azimuth_deg"], npp)
    lift_class("models/prompt_processors/base.py", "DirectionConfig", npp, keep_fields=True)
    lift("models/prompt_processors/base.py", ["configure"], npp, cls="PromptProcessor")
    lift("models/prompt_processors/base.py", ["get_text_embeddings"], npp, cls="PromptProcessorOutput")
    pp = Fake(cfg=Fake(view_dependent_prompt_front=False, front_threshold=45.0, back_threshold=45.0, overhead_threshold=60.0)).bind(npp, ["configure"])
    try:
        pp.configure()            # builds self.directions / direction2idx (:243-309), then reaches for load/prompt_library.json
    except (FileNotFoundError, OSError):
        pass
    assert [d.name for d in pp.directions] == ["side", "front", "back", "overhead"]
    Nv, Nt, Nf = 4, 5, 3
    vd = torch.arange(Nv, dtype=torch.float32).view(Nv, 1, 1).expand(Nv, Nt, Nf).contiguous()
    po = Fake(text_embeddings=vd[:1] + 100, uncond_text_embeddings=vd[:1] + 200, null_text_embeddings=vd[:1] + 300, text_embeddings_vd=vd + 10,
              uncond_text_embeddings_vd=vd + 20, directions=pp.directions, direction2idx=pp.direction2idx).bind(npp, ["get_text_embeddings"])
    el = torch.tensor([0.0, 0.0, 0.0, 0.0, 70.0, 0.0, 61.0, 10.0, -15.0, 59.9])
    az = torch.tensor([90.0, 10.0, -44.0, 170.0, 10.0, -136.0, 179.0, 45.0, 225.0, 315.1])
    dist = torch.ones_like(el)
    G["prompt"] = {"elevation": el, "azimuth": az, "vd": po.get_text_embeddings(el, az, dist, True, return_null_text_embeddings=True),
                   "no_vd": po.get_text_embeddings(el, az, dist, False, return_null_text_embeddings=True),
                   "tables": {"text": po.text_embeddings, "uncond": po.uncond_text_embeddings, "null": po.null_text_embeddings,
                              "text_vd": po.text_embeddings_vd, "uncond_vd": po.uncond_text_embeddings_vd}}
    # branch layout of the noise prediction (compute_without_perpneg :388-438): latents x3, t x3, chunk(3) = text | uncond | null
    lift("models/guidance/dreammat_guidance.py", ["compute_without_perpneg"], ng, cls="StableDiffusionLightGuidance")
    import contextlib
    rec = {}

    def fake_unet(unet, x, t_, encoder_hidden_states=None, **kw):
        rec.update(x=x.clone(), t=t_.clone(), ctx=encoder_hidden_states.clone())
        return x * 2 + encoder_hidden_states.mean(dim=(1, 2)).view(-1, 1, 1, 1)
    gw = Fake(cfg=Fake(view_dependent_prompting=True), unet=None, forward_unet=fake_unet,
              disable_unet_class_embedding=lambda unet: contextlib.nullcontext(unet)).bind(ng, ["compute_without_perpneg"])
    Bp = 3
    zn = torch.randn(Bp, 4, 2, 2, generator=g)
    tt = torch.tensor([5, 400, 900])
    e_t, e_u, e_n = gw.compute_without_perpneg([0], po, zn, tt, el[:Bp], az[:Bp], dist[:Bp], [])
    G["branches"] = {"latents_noisy": zn, "t": tt, "unet_in": rec["x"], "unet_t": rec["t"], "unet_ctx": rec["ctx"], "eps_text": e_t,
                     "eps_uncond": e_u, "eps_null": e_n}

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

    # ------------------------------------------------------------------ mesh normalisation block (dreammat_mesh.py:163-197)
    nmn = base_ns()
    span = lift_block("models/geometry/dreammat_mesh.py", "centroid = mesh.vertices.mean(0)", "mesh.vertices = np.dot(mesh2std, mesh.vertices.T).T",
                      nmn, "normalize_block", ["self", "mesh"], "mesh.vertices")
    verts = np.random.RandomState(7).randn(40, 3) * np.array([1.0, 2.5, 0.7]) + np.array([0.3, -1.0, 2.0])
    cases_m = []
    for up, front, sc in (("+z", "+x", 0.8), ("+y", "+z", 0.5), ("-y", "+x", 0.8), ("+x", "-z", 1.0)):
        me = Fake(cfg=Fake(shape_init_mesh_up=up, shape_init_mesh_front=front, shape_init_params=sc)).bind(nmn, ["normalize_block"])
        cases_m.append((up, front, sc, torch.from_numpy(me.normalize_block(Fake(vertices=verts.copy())))))
    G["mesh_normalize"] = {"vertices": torch.from_numpy(verts), "cases": cases_m, "lines": span}

    # ------------------------------------------------------------------ feature MLP (a3): module structure + forward
    nn_ = base_ns(); nn_["get_activation"] = ns["get_activation"]
    MLP = lift_class("models/networks.py", "VanillaMLP", nn_)
    torch.manual_seed(5)
    mlp = MLP(32, 5, {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64, "n_hidden_layers": 1})
    enc = torch.randn(21, 32, generator=g)
    G["mlp"] = {"state_dict": {k: v.detach().clone() for k, v in mlp.state_dict().items()}, "enc": enc, "out": mlp(enc).detach()}

    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; keys:", sorted(G))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: golden vectors can only be regenerated where /root/reference exists")
    main()
