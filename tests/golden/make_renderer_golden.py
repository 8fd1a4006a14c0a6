"""Golden vectors for the renderer's composition (a2 / a3 / a6), made by EXECUTING the reference's own
`RaytraceRender.forward` (models/renderers/raytracing_renderer.py:110-222) together with its `NVDiffRasterizerContext`
wrapper (utils/rasterize.py:7-78), `get_orthogonal_directions`, `compute_controlnet_normals`, `xfm_vectors`, and -- as the
material -- the reference's own `DreamMatMaterial.forward` / `shade_raytracing` (as in make_golden.py's "material" entry).

Native dependencies that are absent here, and what stands in for each (stated once; everything else is reference code):

  nvdiffrast `dr.rasterize`    -> the oracle's pixel-centre ray cast (oracle.render.gbuffer), returned in nvdiffrast's layout
                                  (u, v, z/w, triangle id + 1)
  nvdiffrast `dr.interpolate`  -> u * a[t0] + v * a[t1] + (1 - u - v) * a[t2] from that `rast`
  nvdiffrast `dr.antialias`    -> the oracle's silhouette blend (antialias_pairs / antialias_apply)
  tiny-cuda-nn geometry        -> oracle.render.geometry_forward on a small hash grid stored with the vectors
  `_raytracing` BVH            -> an analytic sphere occluder (make_golden.sphere_tracer)

so the vectors pin what the REFERENCE wrote around those calls: mask / selector, the inverse-depth map and its min / max
normalisation, the view-space normal map over its (0.5, 0.5, 1) background, the tangent-frame jitter and the order of its two
random draws, the canvases of ones with the covered pixels written in, which outputs are antialiased, and the twelve output
keys and shapes.  `.to("cuda")` / `.to(self.device)` are redirected to the CPU while the function runs and, like a real host ->
device transfer, return a new tensor for a leaf that requires grad (the source is not edited).

Run from the repo root where /root/reference exists:  python tests/golden/make_renderer_golden.py -> renderer_vectors.pt
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import render as O                                                            # noqa: E402
from tests._fixtures import make_scene                                                    # noqa: E402
from tests.golden.make_golden import REF, Fake, base_ns, lift, lift_class, sphere_tracer  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "renderer_vectors.pt")


def main():
    sc = make_scene(res=20, subdiv=2, bump=0.1, seed=4, n_views=1)
    gb, cam = sc["gb"], sc["cam"]
    H = W = sc["res"]
    v, f, vn = sc["v"], sc["f"], sc["vn"]
    pairs = O.antialias_pairs(gb["rast"][0], v, f, cam["mvp_mtx"][0])
    seen = {}

    class Dr:
        class RasterizeGLContext:
            def __init__(self, device=None):
                pass

        @staticmethod
        def rasterize(ctx, pos, tri, resolution, grad_db=True):
            seen["clip"] = pos.clone()
            assert tuple(resolution) == (H, W) and tri.dtype == torch.int32
            return gb["rast"].clone(), None

        @staticmethod
        def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
            t = (rast[..., 3].long() - 1)
            tt = tri.long()[t.clamp(min=0)]
            u, w_ = rast[..., 0:1], rast[..., 1:2]
            a = attr[0]
            out = u * a[tt[..., 0]] + w_ * a[tt[..., 1]] + (1 - u - w_) * a[tt[..., 2]]
            return torch.where((t >= 0)[..., None], out, torch.zeros_like(out)), None

        @staticmethod
        def antialias(color, rast, pos, tri):
            B, h, w, c = color.shape
            return torch.stack([O.antialias_apply(color[b].reshape(h * w, c), pairs).reshape(h, w, c) for b in range(B)])

    # --- the reference's rasteriser wrapper and renderer methods
    nr = base_ns()
    nr.update(dr=Dr, Union=None, Tuple=None)
    Ctx = lift_class("utils/rasterize.py", "NVDiffRasterizerContext", nr)
    ns = base_ns()
    lift("models/renderers/raytracing_renderer.py", ["xfm_vectors"], ns)
    lift("models/renderers/raytracing_renderer.py", ["forward", "get_orthogonal_directions", "compute_controlnet_normals"], ns, cls="RaytraceRender")
    # --- the reference's material (ray-traced branch), as in make_golden.py
    no = base_ns()
    lift("utils/ops.py", ["dot", "reflect", "scale_tensor", "get_activation"], no)
    nm = base_ns()
    nm["get_activation"] = no["get_activation"]
    lift("models/materials/dreammat_material.py", ["saturate_dot", "sample_sphere", "material_smoothness_grad"], nm)
    names = ["get_envirmentlight_blender", "get_lights", "fresnel_schlick", "fresnel_schlick_directions", "geometry_schlick_ggx",
             "geometry_schlick", "get_orthogonal_directions", "sample_diffuse_directions", "sample_specular_directions",
             "distribution_ggx", "geometry", "shade_raytracing", "forward", "set_raytracer"]
    lift("models/materials/dreammat_material.py", names, nm, cls="DreamMatMaterial")
    ND, NS = 24, 16
    g = torch.Generator().manual_seed(31)
    env = torch.rand(16, 32, 3, generator=g) * 2.0
    mcfg = Fake(use_raytracing=True, material_activation="sigmoid", min_metallic=0.0, max_metallic=0.9, min_roughness_squre=0.01,
                max_roughness_squre=0.9, min_roughness=0.1, max_roughness=0.95, random_azimuth=True, geometry_type="schlick",
                use_bump=False, diffuse_sample_num=ND, specular_sample_num=NS)
    mat = Fake(cfg=mcfg, light=[env])
    for name, n in (("diffuse_direction_samples", ND), ("specular_direction_samples", NS)):
        az, el = nm["sample_sphere"](n, 0)
        az, el = az * 0.5 / np.pi, 1 - 2 * el / np.pi
        setattr(mat, name, torch.from_numpy(np.stack([az, el], -1).astype(np.float32)))
    mat.bind(nm, names)
    occ = {"center": [0.5, 0.2, 1.1], "radius": 0.5}
    mat.set_raytracer(sphere_tracer(occ["center"], occ["radius"]))
    mat_call = lambda *a, **k: mat.forward(*a, **k)                   # noqa: E731   (`self.material(...)`)
    # --- geometry stand-in: the oracle's hash grid + MLP, small enough to store
    meta, n_entries = O.hashgrid_meta(n_levels=4, log2_T=9)
    grid = (torch.rand(n_entries * 2, generator=g) * 2 - 1) * 0.5
    W1 = torch.randn(64, 8, generator=g) * 0.5
    W2 = torch.randn(5, 64, generator=g) * 0.5

    class Geo:
        cfg = Fake(n_input_dims=3)

        def __call__(self, points, output_normal=False):
            return {"features": O.geometry_forward(points, grid, W1, W2, meta)}

    ren = Fake(device="cpu", ctx=Ctx("gl", "cpu"), mesh=Fake(v_pos=v, t_pos_idx=f, v_nrm=vn), geometry=Geo(), material=mat_call,
               change_type="gaussian", change_eps=0.05)
    ren.bind(ns, ["forward", "get_orthogonal_directions", "compute_controlnet_normals"])
    pn = int(gb["selector"].sum())
    SEED = 4242
    real_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        r = real_to(self, *a, **k)
        # a host -> device transfer yields a NEW (non-leaf) tensor; :187 relies on that for its in-place write into `color`
        return r.clone() if (r is self and self.requires_grad and self.is_leaf) else r
    torch.Tensor.to = to_cpu
    try:
        torch.manual_seed(SEED)
        out = ren.forward(0, cam["rays_o"], cam["rays_d"], cam["w2c"], cam["mvp_mtx"], None, None, H, W)
    finally:
        torch.Tensor.to = real_to
    torch.manual_seed(SEED)      # the four draws of the call, in order: jitter angle (:163), jitter length (:167), shading (:566, :589)
    rand_ang = torch.rand(pn, 1)
    normal_eps = torch.normal(mean=0.0, std=0.05, size=[pn, 1])
    rd = torch.rand((pn, 1, 1)); rs = torch.rand((pn, 1, 1))
    G = {"in": {"v": v, "f": f, "vn": vn, "rays_o": cam["rays_o"], "rays_d": cam["rays_d"], "w2c": cam["w2c"], "mvp_mtx": cam["mvp_mtx"],
                "res": H, "env": env, "occluder": occ, "n_diffuse": ND, "n_specular": NS, "tab_d": mat.diffuse_direction_samples,
                "tab_s": mat.specular_direction_samples, "grid": grid, "W1": W1, "W2": W2, "hash_levels": 4, "hash_log2_T": 9,
                "rand_ang": rand_ang, "normal_eps": normal_eps, "rand_d": rd, "rand_s": rs},
         "standins": {"rast": gb["rast"], "clip_seen_by_rasterize": seen["clip"], "n_aa_pairs": len(pairs)},
         "out": {k: (v_.detach() if torch.is_tensor(v_) else v_) for k, v_ in out.items()}, "pn": pn}
    torch.save(G, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; covered px", pn, "aa pairs", len(pairs), "keys", sorted(out))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present")
    main()
