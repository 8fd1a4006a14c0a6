"""Oracle: render half of the SDS iteration (rows a1-a6 of SURVEY.md section 8).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch-CPU fp32, autograd on, so the
oracle also supplies reference gradients.  Every random draw is an explicit argument
 so the oracle and the CUDA kernels consume identical randomness.

All citations are relative to /root/reference/threestudio_dreammat/threestudio/.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import build as _build

# ----------------------------------------------------------------------------- ray tracing


class RayTracer:
    """CPU stand-in for `_raytracing` as wrapped at models/renderers/raytracing_renderer.py:20-67.

    trace(o, d) -> (positions, face_normals, depth): depth = RT_MAX_DIST (10) on a miss,
    which is what raytracing_renderer.py:322 (`depth >= 10`) keys on.
    """

    def __init__(self, vertices, triangles, brute: bool = False):
        lib = ctypes.CDLL(_build.build())
        self.lib = lib
        self.v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32))
        self.t = np.ascontiguousarray(np.asarray(triangles, dtype=np.int32))
        self.brute = brute
        lib.rt_bvh_build.restype = ctypes.c_void_p
        lib.rt_bvh_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.rt_bvh_free.argtypes = [ctypes.c_void_p]
        lib.rt_trace_bvh.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] + [ctypes.c_void_p] * 3
        lib.rt_trace_brute.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 3
        self.h = None if brute else lib.rt_bvh_build(self.v.ctypes.data, self.t.ctypes.data, len(self.t))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.rt_bvh_free(self.h)
            self.h = None

    def trace_raw(self, ro, rd):
        ro = np.ascontiguousarray(np.asarray(ro, dtype=np.float32).reshape(-1, 3))
        rd = np.ascontiguousarray(np.asarray(rd, dtype=np.float32).reshape(-1, 3))
        n = ro.shape[0]
        t = np.empty(n, np.float32)
        tri = np.empty(n, np.int32)
        uv = np.empty((n, 2), np.float32)
        if self.brute:
            self.lib.rt_trace_brute(self.v.ctypes.data, self.t.ctypes.data, len(self.t), ro.ctypes.data,
                                    rd.ctypes.data, n, t.ctypes.data, tri.ctypes.data, uv.ctypes.data)
        else:
            self.lib.rt_trace_bvh(self.h, ro.ctypes.data, rd.ctypes.data, n, t.ctypes.data, tri.ctypes.data,
                                  uv.ctypes.data)
        return t, tri, uv

    def trace(self, rays_o: torch.Tensor, rays_d: torch.Tensor):
        """raytracing_renderer.py:318-324 `RaytraceRender.trace`: returns hit mask only
        (positions / normals are never consumed by dreammat_material.py:490-507)."""
        t, tri, _ = self.trace_raw(rays_o.detach().numpy(), rays_d.detach().numpy())
        depth = torch.from_numpy(t)
        hit = ~(depth >= 10)
        return depth, hit


# ----------------------------------------------------------------------------- small helpers


def lin2srgb(x):
    """utils/ops.py:83-88."""
    return torch.where(x > 0.0031308, torch.pow(torch.clamp(x, min=0.0031308), 1.0 / 2.4) * 1.055 - 0.055,
                       12.92 * x).clamp(0.0, 1.0)


def saturate_dot(a, b):
    """models/materials/dreammat_material.py:62-63."""
    return torch.clamp(torch.sum(a * b, dim=-1, keepdim=True), min=0.0, max=1.0)


def get_orthogonal_directions(d):
    """dreammat_material.py:542-552 == raytracing_renderer.py:306-316."""
    x, y, z = torch.split(d, 1, dim=-1)
    o0 = torch.cat([y, -x, torch.zeros_like(x)], -1)
    o1 = torch.cat([-z, torch.zeros_like(x), x], -1)
    m0 = torch.norm(o0, dim=-1) > torch.norm(o1, dim=-1)
    o = torch.where(m0[..., None], o0, o1)
    return F.normalize(o, dim=-1)


def sample_sphere(num_samples, begin_elevation=0):
    """dreammat_material.py:89-102 (numpy float64, as in the reference)."""
    ratio = (begin_elevation + 90) / 180
    num_points = int(num_samples // (1 - ratio))
    phi = (np.sqrt(5) - 1.0) / 2.0
    az, el = [], []
    for n in range(num_points - num_samples, num_points):
        z = 2.0 * n / num_points - 1.0
        az.append(2 * np.pi * n * phi % (2 * np.pi))
        el.append(np.arcsin(z))
    return np.array(az), np.array(el)


def direction_tables(n):
    """dreammat_material.py:389-398: (ua, ue) in [0,1], float32, shape [n,2]."""
    az, el = sample_sphere(n, 0)
    az, el = az * 0.5 / np.pi, 1 - 2 * el / np.pi
    return torch.from_numpy(np.stack([az, el], -1).astype(np.float32))


def envmap_lookup(light, directions):
    """dreammat_material.py:439-455 `get_envirmentlight_blender`; light is [H,W,3]."""
    height, width, _ = light.shape
    directions = directions / directions.norm(p=2, dim=-1, keepdim=True)
    x, y, z = directions.unbind(-1)
    theta = torch.acos(z)
    phi = torch.atan2(y, x) % (2 * np.pi)
    u = -phi / (2 * np.pi) + 0.5
    v = theta / np.pi
    xx = (u * width) % width
    yy = (v * height) % height
    return light[yy.long(), xx.long(), :]


# ----------------------------------------------------------------------------- hash grid + MLP (a3)

HG_PRIMES = (1, 2654435761, 805459861)


def hashgrid_meta(n_levels=16, log2_T=19, base=16, scale=1.447269237440378):
    """tiny-cuda-nn GridEncoding level layout (un-vendored dep, requirements.txt:6; call site
    models/networks.py:55-64).  Returns per-level (scale, res, n_entries, offset, hashed)."""
    meta, off = [], 0
    for l in range(n_levels):
        # tcnn: scale = exp2f(level * log2f(per_level_scale)) * base_resolution - 1.0f  (all fp32)
        s = np.float32(np.exp2(np.float32(l) * np.float32(np.log2(np.float32(scale)))) * np.float32(base) - np.float32(1.0))
        res = int(math.ceil(float(s))) + 1
        n = res ** 3
        n = ((n + 7) // 8) * 8
        n = min(n, 1 << log2_T)
        meta.append(dict(scale=float(s), res=res, size=n, offset=off, hashed=(res ** 3 > n)))
        off += n
    return meta, off


def hashgrid_encode(x01, params, meta, n_feat=2):
    """x01 [N,3] in [0,1]; params flat [total*n_feat] -> [N, L*n_feat]; differentiable in params.
    tcnn `kernel_grid` (Hash grid, Linear interpolation): pos = x*scale + 0.5; corner
    index = dense x + y*res + z*res^2 when the level is not hashed, else
    (x*1 ^ y*2654435761 ^ z*805459861) mod size (uint32 arithmetic)."""
    outs = []
    N = x01.shape[0]
    p = params.view(-1, n_feat)
    for m in meta:
        pos = x01 * np.float32(m["scale"]) + 0.5
        pf = torch.floor(pos)
        w = pos - pf
        pi = pf.to(torch.int64)
        acc = torch.zeros(N, n_feat, dtype=params.dtype)
        for c in range(8):
            o = [(c >> k) & 1 for k in range(3)]
            cc = [pi[:, k] + o[k] for k in range(3)]
            ww = torch.ones(N, dtype=x01.dtype)
            for k in range(3):
                ww = ww * (w[:, k] if o[k] else (1.0 - w[:, k]))
            if m["hashed"]:
                idx = torch.zeros(N, dtype=torch.int64)
                for k in range(3):
                    idx = idx ^ ((cc[k] & 0xFFFFFFFF) * HG_PRIMES[k] & 0xFFFFFFFF)
                idx = idx % m["size"]
            else:
                res = m["res"]
                idx = ((cc[0] & 0xFFFFFFFF) + (cc[1] & 0xFFFFFFFF) * res + (cc[2] & 0xFFFFFFFF) * res * res) & 0xFFFFFFFF
                idx = idx % m["size"]
            acc = acc + ww[:, None] * p[m["offset"] + idx]
        outs.append(acc)
    return torch.cat(outs, -1)


def geometry_forward(points, params, W1, W2, meta):
    """models/geometry/dreammat_mesh.py:239-254 with radius 1 bbox (geometry/base.py:20-32,
    utils/ops.py:26-37) and VanillaMLP bias-free 32->64->5 (models/networks.py:150-187)."""
    x01 = (points - (-1.0)) / (1.0 - (-1.0))
    enc = hashgrid_encode(x01, params, meta)
    return mlp_forward(enc, W1, W2)


def mlp_forward(enc, W1, W2):
    """VanillaMLP (models/networks.py:150-187) as configured by dreammat.yaml: Linear(bias=False) -> ReLU ->
    Linear(bias=False), output activation none; W1 = layers.0.weight [64,32], W2 = layers.2.weight [5,64]."""
    return torch.relu(enc @ W1.t()) @ W2.t()


# ----------------------------------------------------------------------------- material (a4)


def material_smoothness_grad(material, material_jitter):
    """dreammat_material.py:110-123."""
    kd = torch.abs(material[..., :3] - material_jitter[..., :3])
    ks = torch.abs(material[..., 3:5] - material_jitter[..., 3:5])
    luma = (kd[..., 0] + kd[..., 1] + kd[..., 2]) / 3
    loss = torch.mean(luma * kd[..., -1]) * 0.25
    loss = loss + torch.mean(ks[..., :-1] * ks[..., -1:]) * 0.1
    return loss


def material_params(features, features_jitter, use_raytracing=True, min_metallic=0.0, max_metallic=0.9,
                    min_r2=0.01, max_r2=0.9, min_r=0.1, max_r=0.95):
    """dreammat_material.py:727-762."""
    m = torch.sigmoid(features)
    mj = torch.sigmoid(features_jitter)
    reg = material_smoothness_grad(m, mj)
    albedo = m[..., :3].clamp(0.0, 1.0)
    metallic = m[..., 3:4] * (max_metallic - min_metallic) + min_metallic
    if use_raytracing:
        rough = m[..., 4:5] * (max_r2 - min_r2) + min_r2
    else:
        rough = m[..., 4:5] * (max_r - min_r) + min_r
    return albedo, metallic, rough, reg


def distribution_ggx(NoH, a):
    """dreammat_material.py:599-604."""
    a2 = a ** 2
    denom = NoH ** 2 * (a2 - 1.0) + 1.0
    return a2 / (np.pi * denom ** 2 + 1e-4)


def geometry_schlick_ggx(NoV, a):
    """dreammat_material.py:519-525."""
    k = a / 2
    return NoV / (NoV * (1 - k) + k + 1e-5)


def shade_raytracing(pts, normals, view_dirs, light, metallic, roughness, albedo, rand_d, rand_s, trace_fn,
                     n_diffuse=200, n_specular=128):
    """dreammat_material.py:615-677 with sample_diffuse_directions :554-573,
    sample_specular_directions :575-596, get_lights :490-507 (is_train=True, random_azimuth=True).

    rand_d / rand_s: [pn,1,1] uniform draws (appendix B #5/#6).  trace_fn(o,d)->hit mask."""
    tab_d = direction_tables(n_diffuse).to(pts.device)      # device-generic: bench.py's stock-PyTorch-CUDA leg runs this on the GPU
    tab_s = direction_tables(n_specular).to(pts.device)
    reflections = torch.sum(view_dirs * normals, -1, keepdim=True) * normals * 2 - view_dirs
    F0 = 0.04 * (1 - metallic) + metallic * albedo

    # diffuse directions
    z = normals
    x = get_orthogonal_directions(normals)
    y = torch.cross(z, x, dim=-1)
    az, el = torch.split(tab_d, 1, dim=1)
    el, az = el.unsqueeze(0), az.unsqueeze(0)
    az = az * torch.pi * 2
    el_sqrt = torch.sqrt(el + 1e-7)
    az = (az + rand_d * torch.pi * 2) % (2 * torch.pi)
    cz = torch.sqrt(1 - el + 1e-7)
    cx = el_sqrt * torch.cos(az)
    cy = el_sqrt * torch.sin(az)
    diffuse_directions = cx * x.unsqueeze(1) + cy * y.unsqueeze(1) + cz * z.unsqueeze(1)

    # specular directions
    z = reflections
    x = get_orthogonal_directions(reflections)
    y = torch.cross(z, x, dim=-1)
    a = roughness
    az, el = torch.split(tab_s, 1, dim=1)
    phi = np.pi * 2 * az
    a_, el = a.unsqueeze(1), el.unsqueeze(0)
    cos_theta = torch.sqrt((1.0 - el + 1e-6) / (1.0 + (a_ ** 2 - 1.0) * el + 1e-6) + 1e-6)
    sin_theta = torch.sqrt(1 - cos_theta ** 2 + 1e-6)
    phi = phi.unsqueeze(0)
    phi = (phi + rand_s * np.pi * 2) % (2 * np.pi)
    cx = torch.cos(phi) * sin_theta
    cy = torch.sin(phi) * sin_theta
    cz = cos_theta
    specular_directions = cx * x.unsqueeze(1) + cy * y.unsqueeze(1) + cz * z.unsqueeze(1)

    dn, sn_ = n_diffuse, n_specular
    NoL_d = saturate_dot(diffuse_directions, normals.unsqueeze(1))
    p_d = NoL_d / np.pi * (dn / (sn_ + dn))
    H_s = F.normalize(view_dirs.unsqueeze(1) + specular_directions, dim=-1)
    NoH_s = saturate_dot(normals.unsqueeze(1), H_s)
    VoH_s = saturate_dot(view_dirs.unsqueeze(1), H_s)
    p_s = distribution_ggx(NoH_s, roughness.unsqueeze(1)) * NoH_s / (4 * VoH_s + 1e-5) * (sn_ / (sn_ + dn))

    directions = torch.cat([diffuse_directions, specular_directions], 1)
    probability = torch.cat([p_d, p_s], 1)
    sn = dn + sn_

    H = F.normalize(view_dirs.unsqueeze(1) + directions, dim=-1)
    HoV = torch.clamp(torch.sum(H * view_dirs.unsqueeze(1), dim=-1, keepdim=True), min=0.0, max=1.0)
    fresnel = F0.unsqueeze(1) + (1.0 - F0.unsqueeze(1)) * torch.clamp(1.0 - HoV, min=0.0, max=1.0) ** 5.0
    NoV = saturate_dot(normals, view_dirs).unsqueeze(1)
    NoL = saturate_dot(normals.unsqueeze(1), directions)
    geometry = geometry_schlick_ggx(NoV, roughness.unsqueeze(1)) * geometry_schlick_ggx(NoL, roughness.unsqueeze(1))
    NoH = saturate_dot(normals.unsqueeze(1), H)
    distribution = distribution_ggx(NoH, roughness.unsqueeze(1))

    # get_lights (:490-507)
    pts_ = pts.unsqueeze(1).repeat(1, sn, 1)
    o = pts_.reshape(-1, 3) + directions.reshape(-1, 3).detach() * 1e-5
    hit = trace_fn(o, directions.reshape(-1, 3).detach()).reshape(-1, sn)
    lights = torch.zeros(pts.shape[0], sn, 3, dtype=light.dtype, device=pts.device)
    miss = ~hit
    if miss.any():
        lights[miss] = envmap_lookup(light, directions.detach()[miss])

    w = distribution * geometry / (4 * NoV * probability + 1e-5)
    specular_lights = lights * w
    specular_colors = torch.mean(fresnel * specular_lights, 1)
    diffuse_lights = lights[:, :dn]
    diffuse_colors = torch.mean(albedo.unsqueeze(1) * diffuse_lights, 1)
    colors = lin2srgb(diffuse_colors + specular_colors)
    out = {
        "color": colors,
        "albedo": lin2srgb(albedo.detach()),
        "roughness": torch.sqrt(roughness + 1e-7).detach(),
        "metalness": metallic.detach(),
        "specular_lights": lin2srgb(torch.mean(lights[:, dn:, :], dim=1)),
        "diffuse_lights": lin2srgb(torch.mean(lights[:, :dn, :], dim=1)),
        "specular_colors": lin2srgb(specular_colors.detach()),
        "diffuse_colors": lin2srgb(diffuse_colors.detach()),
        "_hit": hit,
        "_color_linear": (diffuse_colors + specular_colors).detach(),
    }
    return out


# ----------------------------------------------------------------------------- split-sum (a5)

CUBE_RES_DEFAULT = 128  # envlight default max_res (un-vendored)


def cube_to_dir(s, x, y):
    """envlight / nvdiffrec util.cube_to_dir (un-vendored dep, requirements.txt:24)."""
    one = torch.ones_like(x)
    if s == 0:
        r = (one, -y, -x)
    elif s == 1:
        r = (-one, -y, x)
    elif s == 2:
        r = (x, one, y)
    elif s == 3:
        r = (x, -one, -y)
    elif s == 4:
        r = (x, -y, one)
    else:
        r = (-x, -y, -one)
    return torch.stack(r, dim=-1)


def _tex2d_linear_wrapclamp(img, uv):
    """nvdiffrast dr.texture(filter='linear', boundary='wrap') on a 2-D map; img [H,W,C], uv [...,2]."""
    H, W, _ = img.shape
    x = uv[..., 0] * W - 0.5
    y = uv[..., 1] * H - 0.5
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    fx = (x - x0)[..., None]
    fy = (y - y0)[..., None]
    x0 = x0.long()
    y0 = y0.long()
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = x0 % W, x1 % W
    y0, y1 = y0 % H, y1 % H
    return (img[y0, x0] * (1 - fx) * (1 - fy) + img[y0, x1] * fx * (1 - fy) + img[y1, x0] * (1 - fx) * fy +
            img[y1, x1] * fx * fy)


def latlong_to_cubemap(latlong, res):
    """envlight utils.latlong_to_cubemap (restated)."""
    cube = torch.zeros(6, res, res, latlong.shape[-1])
    lin = torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    for s in range(6):
        v = F.normalize(cube_to_dir(s, gx, gy), dim=-1)
        tu = torch.atan2(v[..., 0:1], -v[..., 2:3]) / (2 * np.pi) + 0.5
        tv = torch.acos(torch.clamp(v[..., 1:2], min=-1, max=1)) / np.pi
        cube[s] = _tex2d_linear_wrapclamp(latlong, torch.cat((tu, tv), dim=-1))
    return cube


def _texel_dirs_area(res):
    lin = (torch.arange(res, dtype=torch.float64) + 0.5) / res * 2 - 1
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    dirs = torch.stack([F.normalize(cube_to_dir(s, gx, gy), dim=-1) for s in range(6)])  # [6,r,r,3]

    def area(x, y):
        return torch.atan2(x * y, torch.sqrt(x * x + y * y + 1))
    h = 1.0 / res
    x0, x1, y0, y1 = gx - h, gx + h, gy - h, gy + h
    sa = area(x0, y0) - area(x0, y1) - area(x1, y0) + area(x1, y1)
    return dirs.float(), sa.float().unsqueeze(0).expand(6, -1, -1)


def diffuse_cubemap(cube):
    """nvdiffrec renderutils.diffuse_cubemap (restated): cosine-weighted convolution,
    cos clamped to [0, 0.999], weight = cos * texel solid angle / pi."""
    res = cube.shape[1]
    dirs, sa = _texel_dirs_area(res)
    D = dirs.reshape(-1, 3)
    Lm = cube.reshape(-1, 3)
    w = torch.clamp(D @ D.t(), 0.0, 0.999) * sa.reshape(1, -1) / 3.141592
    return (w @ Lm).reshape(6, res, res, 3)


def ndf_cutoff(roughness, cutoff=0.99, n=1000000):
    """nvdiffrec renderutils.__ndfBounds (restated): cos of the half-angle inside which the
    GGX NDF (alpha^2 = roughness^4) holds `cutoff` of its cumulative mass."""
    ct = np.cos(np.linspace(0, np.pi / 2.0, n))
    a2 = roughness ** 4
    d = (ct * a2 - ct) * ct + 1.0
    Dn = np.cumsum(a2 / (d * d * np.pi))
    idx = np.argmax(Dn >= Dn[-1] * cutoff)
    return float(ct[idx])


def specular_cubemap(cube, roughness, cutoff=0.99):
    """nvdiffrec renderutils.specular_cubemap (restated): V=N=R GGX prefilter, weight =
    NoL * D(alpha^2, N.H) * solid_angle / 4 over texels with L.N >= costheta_cutoff."""
    res = cube.shape[1]
    dirs, sa = _texel_dirs_area(res)
    D = dirs.reshape(-1, 3)
    Lm = cube.reshape(-1, 3)
    cc = ndf_cutoff(roughness, cutoff)
    a2 = (roughness * roughness) ** 2
    out = torch.zeros_like(Lm)
    chunk = 2048
    for i in range(0, D.shape[0], chunk):
        N = D[i:i + chunk]
        LdN = N @ D.t()
        Hh = F.normalize(N[:, None, :] + D[None, :, :], dim=-1)
        NoH = torch.clamp((Hh * N[:, None, :]).sum(-1), min=0.0)
        dd = (NoH * a2 - NoH) * NoH + 1.0
        ndf = a2 / (dd * dd * np.pi)
        w = torch.clamp(LdN, min=0.0) * ndf * sa.reshape(1, -1) / 4.0
        w = torch.where(LdN >= cc, w, torch.zeros_like(w))
        out[i:i + chunk] = (w @ Lm) / w.sum(-1, keepdim=True)
    return out.reshape(6, res, res, 3)


def build_envlight(latlong_hdr, scale=2.0, max_res=CUBE_RES_DEFAULT, min_res=16, min_rough=0.08, max_rough=0.5):
    """envlight.EnvLight.__init__/build_mips (restated).  Returns (diffuse [6,16,16,3], [specular mips])."""
    base = latlong_to_cubemap(latlong_hdr * scale, max_res)
    spec = [base]
    while spec[-1].shape[1] > min_res:
        c = spec[-1]
        spec.append(F.avg_pool2d(c.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous())
    diffuse = diffuse_cubemap(spec[-1])
    for i in range(len(spec) - 1):
        r = (i / max(len(spec) - 2, 1)) * (max_rough - min_rough) + min_rough
        spec[i] = specular_cubemap(spec[i], r)
    spec[-1] = specular_cubemap(spec[-1], 1.0)
    return diffuse, spec


def envlight_mip_level(roughness, n_mips, min_rough=0.08, max_rough=0.5):
    """envlight.EnvLight.get_mip (restated)."""
    return torch.where(
        roughness < max_rough,
        (torch.clamp(roughness, min_rough, max_rough) - min_rough) / (max_rough - min_rough) * (n_mips - 2),
        (torch.clamp(roughness, max_rough, 1.0) - max_rough) / (1.0 - max_rough) + n_mips - 2)


def dir_to_cube(d):
    """nvdiffrast cube-map face selection (OpenGL convention, inverse of cube_to_dir)."""
    ax = d.abs()
    x, y, z = d.unbind(-1)
    fx = (ax[..., 0] >= ax[..., 1]) & (ax[..., 0] >= ax[..., 2])
    fy = (~fx) & (ax[..., 1] >= ax[..., 2])
    face = torch.where(fx, torch.where(x >= 0, 0, 1), torch.where(fy, torch.where(y >= 0, 2, 3),
                                                                  torch.where(z >= 0, 4, 5)))
    ma = torch.where(fx, ax[..., 0], torch.where(fy, ax[..., 1], ax[..., 2]))
    # s,t such that cube_to_dir(face, s, t) ~ d / ma
    s = torch.where(face == 0, -z, torch.where(face == 1, z, torch.where(face == 5, -x, x))) / ma
    t = torch.where(face == 2, z, torch.where(face == 3, -z, -y)) / ma
    return face, s, t


def _cube_fetch(cube, face, ix, iy):
    """Seamless texel fetch: out-of-face texels come from the adjacent face (re-projected
    texel centre); corner texels (both indices outside) are dropped -> weight 0."""
    res = cube.shape[1]
    inx = (ix >= 0) & (ix < res)
    iny = (iy >= 0) & (iy < res)
    valid = inx | iny
    s = (ix.float() + 0.5) / res * 2 - 1
    t = (iy.float() + 0.5) / res * 2 - 1
    val = torch.zeros(*face.shape, cube.shape[-1])
    for f in range(6):
        m = face == f
        if not m.any():
            continue
        d = cube_to_dir(f, s[m], t[m])
        f2, s2, t2 = dir_to_cube(d)
        jx = torch.clamp(torch.floor((s2 + 1) * 0.5 * res).long(), 0, res - 1)
        jy = torch.clamp(torch.floor((t2 + 1) * 0.5 * res).long(), 0, res - 1)
        val[m] = cube[f2, jy, jx]
    return val, valid


def cube_sample_linear(cube, d):
    """nvdiffrast dr.texture(filter='linear', boundary='cube') (restated)."""
    res = cube.shape[1]
    face, s, t = dir_to_cube(d)
    x = (s + 1) * 0.5 * res - 0.5
    y = (t + 1) * 0.5 * res - 0.5
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    fx, fy = x - x0, y - y0
    x0, y0 = x0.long(), y0.long()
    acc = 0
    wsum = 0
    for dx, dy, w in ((0, 0, (1 - fx) * (1 - fy)), (1, 0, fx * (1 - fy)), (0, 1, (1 - fx) * fy), (1, 1, fx * fy)):
        v, ok = _cube_fetch(cube, face, x0 + dx, y0 + dy)
        w = w * ok.float()
        acc = acc + v * w[..., None]
        wsum = wsum + w
    return acc / wsum[..., None]


def cube_sample_trilinear(mips, d, level):
    """dr.texture(filter='linear-mipmap-linear', mip_level_bias=level, boundary='cube') with an
    explicit mip stack and no uv derivatives: level is clamped to [0, n-1]."""
    n = len(mips)
    lv = torch.clamp(level, 0.0, float(n - 1))
    l0 = torch.floor(lv).long().clamp(max=n - 1)
    l1 = torch.clamp(l0 + 1, max=n - 1)
    f = (lv - l0.float())[..., None]
    out0 = torch.zeros(*d.shape[:-1], 3)
    out1 = torch.zeros(*d.shape[:-1], 3)
    for i in range(n):
        m0 = l0 == i
        if m0.any():
            out0[m0] = cube_sample_linear(mips[i], d[m0])
        m1 = l1 == i
        if m1.any():
            out1[m1] = cube_sample_linear(mips[i], d[m1])
    return out0 * (1 - f) + out1 * f


def fg_lookup(lut, ndv, rough):
    """dr.texture(FG_LUT[1,256,256,2], uv=(ndv,rough), 'linear', 'clamp') -- dreammat_material.py:686-692."""
    H, W, _ = lut.shape
    x = torch.clamp(ndv, 0, 1) * W - 0.5
    y = torch.clamp(rough, 0, 1) * H - 0.5
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    fx, fy = (x - x0)[..., None], (y - y0)[..., None]
    x0, y0 = x0.long(), y0.long()
    x1, y1 = (x0 + 1).clamp(0, W - 1), (y0 + 1).clamp(0, H - 1)
    x0, y0 = x0.clamp(0, W - 1), y0.clamp(0, H - 1)
    return (lut[y0, x0] * (1 - fx) * (1 - fy) + lut[y0, x1] * fx * (1 - fy) + lut[y1, x0] * (1 - fx) * fy +
            lut[y1, x1] * fx * fy)


def shade_splitsum_with(normals, viewdirs, fg_fn, diffuse_fn, specular_fn, metallic, roughness, albedo):
    """dreammat_material.py:679-711 with its three texture fetches as callables: `fg_fn(n.v, roughness) -> [N,2]`
    (dr.texture on FG_LUT), `diffuse_fn(n) -> [N,3]` and `specular_fn(r, roughness) -> [N,3]` (the envlight object).
    Pinned against the reference's own `shade_splitsum` body by tests/golden/make_splitsum_golden.py."""
    v = viewdirs
    n_dot_v = (normals * v).sum(-1, keepdim=True)
    reflective = n_dot_v * normals * 2 - v
    fg = fg_fn(n_dot_v[..., 0], roughness[..., 0])
    F0 = (1 - metallic) * 0.04 + metallic * albedo
    specular_albedo = F0 * fg[:, 0:1] + fg[:, 1:2]
    diffuse_light = diffuse_fn(normals)
    specular_light = specular_fn(reflective, roughness)
    color = (albedo * diffuse_light + specular_albedo * specular_light).clamp(0.0, 1.0)
    return {
        "color": color,
        "albedo": albedo.detach(),
        "roughness": roughness.detach(),
        "metalness": metallic.detach(),
        "specular_lights": lin2srgb(specular_light.detach()),
        "diffuse_lights": lin2srgb(diffuse_light.detach()),
        "specular_colors": lin2srgb(specular_albedo.detach()),
        "diffuse_colors": lin2srgb(albedo.detach()),
    }


def shade_splitsum(normals, viewdirs, diffuse_cube, spec_mips, lut, metallic, roughness, albedo):
    """dreammat_material.py:679-711 over the oracle's own envlight restatement (cube maps) and FG lookup."""
    return shade_splitsum_with(
        normals, viewdirs, lambda ndv, r: fg_lookup(lut, ndv, r), lambda n: cube_sample_linear(diffuse_cube, n),
        lambda d, r: cube_sample_trilinear(spec_mips, d, envlight_mip_level(r[..., 0], len(spec_mips))),
        metallic, roughness, albedo)


# ----------------------------------------------------------------------------- cameras / G-buffer (a1, a2)


def get_projection_matrix(fovy, aspect, near=0.1, far=1000.0):
    """utils/ops.py:266-278 (y flipped for the nvdiffrast convention)."""
    B = fovy.shape[0]
    p = torch.zeros(B, 4, 4)
    p[:, 0, 0] = 1.0 / (torch.tan(fovy / 2.0) * aspect)
    p[:, 1, 1] = -1.0 / torch.tan(fovy / 2.0)
    p[:, 2, 2] = -(far + near) / (far - near)
    p[:, 2, 3] = -2.0 * far * near / (far - near)
    p[:, 3, 2] = -1.0
    return p


def camera_batch(elevation_deg, azimuth_deg, distance, fovy_deg, H, W):
    """data/uncond.py:723-821 camera part + utils/ops.py:179-292 (perturbs = 0, z-up)."""
    el = elevation_deg * math.pi / 180
    az = azimuth_deg * math.pi / 180
    pos = torch.stack([distance * torch.cos(el) * torch.cos(az), distance * torch.cos(el) * torch.sin(az),
                       distance * torch.sin(el)], -1)
    center = torch.zeros_like(pos)
    up = torch.tensor([0.0, 0.0, 1.0])[None].repeat(pos.shape[0], 1)
    lookat = F.normalize(center - pos, dim=-1)
    right = F.normalize(torch.cross(lookat, up, dim=-1), dim=-1)
    up = F.normalize(torch.cross(right, lookat, dim=-1), dim=-1)
    c2w3x4 = torch.cat([torch.stack([right, up, -lookat], dim=-1), pos[:, :, None]], dim=-1)
    c2w = torch.cat([c2w3x4, torch.zeros_like(c2w3x4[:, :1])], dim=1)
    c2w[:, 3, 3] = 1.0
    fovy = fovy_deg * math.pi / 180
    focal = 0.5 * H / torch.tan(0.5 * fovy)
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32) + 0.5, torch.arange(H, dtype=torch.float32) + 0.5,
                          indexing="xy")
    dirs = torch.stack([(i - W / 2), -(j - H / 2), -torch.ones_like(i)], -1)  # unit focal below
    dirs = dirs[None].repeat(pos.shape[0], 1, 1, 1)
    dirs[..., :2] = dirs[..., :2] / focal[:, None, None, None]
    rays_d = (dirs[..., None, :] * c2w[:, None, None, :3, :3]).sum(-1)
    rays_d = F.normalize(rays_d, dim=-1)
    rays_o = c2w[:, None, None, :3, 3].expand(rays_d.shape)
    proj = get_projection_matrix(fovy, W / H)
    w2c = torch.zeros(pos.shape[0], 4, 4)
    w2c[:, :3, :3] = c2w[:, :3, :3].permute(0, 2, 1)
    w2c[:, :3, 3:] = -c2w[:, :3, :3].permute(0, 2, 1) @ c2w[:, :3, 3:]
    w2c[:, 3, 3] = 1.0
    mvp = proj @ w2c
    return dict(rays_o=rays_o.contiguous(), rays_d=rays_d.contiguous(), mvp_mtx=mvp, c2w=c2w, w2c=w2c,
                camera_positions=pos)


def vertex_normals(v, f):
    """models/mesh.py `_compute_vertex_normal` (area-weighted face normals, normalised)."""
    i0, i1, i2 = f[:, 0].long(), f[:, 1].long(), f[:, 2].long()
    fn = torch.cross(v[i1] - v[i0], v[i2] - v[i0], dim=-1)
    vn = torch.zeros_like(v)
    vn.index_add_(0, i0, fn)
    vn.index_add_(0, i1, fn)
    vn.index_add_(0, i2, fn)
    vn = torch.where((vn * vn).sum(-1, keepdim=True) > 1e-20, vn, torch.tensor([0.0, 0.0, 1.0]))
    return F.normalize(vn, dim=-1)


def controlnet_depth(depth, mask, min_val=0.3):
    """raytracing_renderer.py:129-134 / compute_controlnet_depth :333-343: inverse depth of the covered pixels
    rescaled to [min_val, 1] by their min / max, 0 elsewhere."""
    depth = depth.clone()
    dm = 1.0 / (depth[mask] + 1e-6)
    depth[mask] = (1 - min_val) * (dm - dm.min()) / (dm.max() - dm.min() + 1e-6) + min_val
    depth[~mask] = 0.0
    return depth


def controlnet_view_normals(normals, w2c):
    """compute_controlnet_normals (raytracing_renderer.py:326-331) for one view: rotate into the camera frame
    (xfm_vectors :69-83, w = 0), normalise, map to [0,1], flip x."""
    n4 = torch.cat([normals, torch.zeros(normals.shape[0], 1)], -1)
    nv = F.normalize((n4 @ w2c.t())[:, :3], dim=-1)
    nc = 0.5 * (nv + 1)
    nc[..., 0] = 1.0 - nc[..., 0]
    return nc


def gbuffer(tracer: RayTracer, v_pos, t_idx, v_nrm, rays_o, rays_d, mvp, w2c):
    """G-buffer stage of RaytraceRender.forward (raytracing_renderer.py:122-159).

    The reference rasterises with nvdiffrast; visibility at pixel centres is the closest hit
    of the pixel-centre ray (the `rays_d` of data/uncond.py go through i+0.5, j+0.5), so the
    oracle casts those rays.  Returns rast (u, v, z/w, tri_id+1) in nvdiffrast's convention
    (u,v = barycentric weights of vertex 0 and 1), interpolated normal/position, masks, and
    the depth / view-normal maps (without the antialias pass)."""
    B, H, W, _ = rays_d.shape
    t, tri, uv = tracer.trace_raw(rays_o.reshape(-1, 3).numpy(), rays_d.reshape(-1, 3).numpy())
    tri = torch.from_numpy(tri).long()
    hit = tri >= 0
    b1 = torch.from_numpy(uv[:, 0])  # weight of vertex 1
    b2 = torch.from_numpy(uv[:, 1])  # weight of vertex 2
    b0 = 1 - b1 - b2
    tt = t_idx.long()[tri.clamp(min=0)]
    P = v_pos[tt[:, 0]] * b0[:, None] + v_pos[tt[:, 1]] * b1[:, None] + v_pos[tt[:, 2]] * b2[:, None]
    Nn = v_nrm[tt[:, 0]] * b0[:, None] + v_nrm[tt[:, 1]] * b1[:, None] + v_nrm[tt[:, 2]] * b2[:, None]
    P = torch.where(hit[:, None], P, torch.zeros_like(P))
    Nn = torch.where(hit[:, None], Nn, torch.zeros_like(Nn))
    Nn = F.normalize(Nn, dim=-1)
    Ph = torch.cat([P, torch.ones_like(P[:, :1])], -1).reshape(B, H * W, 4)
    clip = (Ph @ mvp.transpose(1, 2)).reshape(-1, 4)
    zw = clip[:, 2] / clip[:, 3]
    rast = torch.stack([b0, b1, zw, (tri + 1).float()], -1)
    rast = torch.where(hit[:, None], rast, torch.zeros_like(rast)).reshape(B, H, W, 4)
    mask = rast[..., 3:] > 0
    # depth normalisation (:129-134), global min / max over the batch
    depth = controlnet_depth(rast[..., 2:3], mask)
    # controlnet view normals (:139-147, :326-331) -- per view so B>1 is well defined (a0)
    nc = torch.stack([controlnet_view_normals(Nn.reshape(B, H * W, 3)[b], w2c[b]) for b in range(B)])
    bg = torch.tensor([0.5, 0.5, 1.0])
    comp_normal = torch.where(mask.reshape(B, H * W, 1), nc, bg).reshape(B, H, W, 3)
    return dict(rast=rast, mask=mask, selector=mask[..., 0].reshape(B, H * W), gb_pos=P.reshape(B, H * W, 3),
                gb_normal=Nn.reshape(B, H * W, 3), comp_depth=depth, comp_normal=comp_normal,
                gb_viewdirs=-rays_d.reshape(B, H * W, 3))


def jitter_positions(positions, normals, rand_ang, normal_eps):
    """raytracing_renderer.py:161-173 ('gaussian'); rand_ang~U[0,1) [pn,1], normal_eps~N(0,0.05) [pn,1]."""
    x = get_orthogonal_directions(normals)
    y = torch.cross(normals, x, dim=-1)
    ang = rand_ang * np.pi * 2
    return positions + (torch.cos(ang) * x + torch.sin(ang) * y) * normal_eps


def render_forward(gb, pairs, geometry_fn, material_fn, rand_ang, normal_eps):
    """RaytraceRender.forward (raytracing_renderer.py:110-222) for ONE view (the reference's inline normal-map code only
    reshapes for B = 1): G-buffer `gb` (gbuffer() above) -> tangent jitter -> geometry twice -> material -> canvases of
    ones with the covered pixels written in -> antialias.  `pairs` = antialias_pairs() of the view; `geometry_fn(points)`
    -> features; `material_fn(pts, features, features_jitter, viewdirs, normals)` -> (shade outputs, mat_reg).
    Pinned against the reference's own forward by tests/golden/make_renderer_golden.py."""
    B, H, W, _ = gb["rast"].shape
    assert B == 1
    sel = gb["selector"][0]
    pos, nrm, vd = gb["gb_pos"][0][sel], gb["gb_normal"][0][sel], gb["gb_viewdirs"][0][sel]
    f = geometry_fn(pos)
    fj = geometry_fn(jitter_positions(pos, nrm, rand_ang, normal_eps))
    shade, reg = material_fn(pos, f, fj, vd, nrm)
    idx = (torch.nonzero(sel).view(-1),)

    def canvas(v):
        return torch.ones(H * W, v.shape[-1], dtype=v.dtype).index_put(idx, v)

    def aa(img):
        return antialias_apply(img.reshape(H * W, -1), pairs).reshape(1, H, W, -1)
    out = {"comp_rgb": aa(canvas(shade["color"])), "opacity": aa(gb["mask"].float()), "comp_depth": gb["comp_depth"],
           "comp_normal": aa(gb["comp_normal"]), "loss_mat_reg": reg}
    for k_out, k_in in (("albedo", "albedo"), ("metalness", "metalness"), ("roughness", "roughness"), ("specular_light", "specular_lights"),
                        ("diffuse_light", "diffuse_lights"), ("specular_color", "specular_colors"), ("diffuse_color", "diffuse_colors")):
        out[k_out] = canvas(shade[k_in].detach()).reshape(1, H, W, -1)
    return out


# ----------------------------------------------------------------------------- procedural fixtures


def icosphere(subdiv=3, radius=0.8, bump=0.0):
    """Test mesh: subdivided icosahedron, optionally displaced so that it self-occludes."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        cache = {}
        vl = list(v)
        nf = []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = (vl[a] + vl[b]) / 2
                vl.append(m / np.linalg.norm(m))
                cache[k] = len(vl) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v = np.array(vl)
        f = np.array(nf)
    if bump:
        r = 1 + bump * (np.sin(5 * v[:, 0]) * np.cos(4 * v[:, 1]) + np.sin(6 * v[:, 2] + 1.0))
        v = v * r[:, None]
    v = v / np.abs(v).max() * radius
    return torch.from_numpy(v.astype(np.float32)), torch.from_numpy(f.astype(np.int32))


def synthetic_envmap(H=256, W=512, seed=0):
    """HDR lat-long test map: sky gradient + a bright sun + low-amplitude noise (max ~ 60)."""
    g = torch.Generator().manual_seed(seed)
    v = (torch.arange(H, dtype=torch.float32) + 0.5) / H
    u = (torch.arange(W, dtype=torch.float32) + 0.5) / W
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    sky = torch.stack([0.4 + 0.3 * (1 - vv), 0.5 + 0.3 * (1 - vv), 0.7 + 0.5 * (1 - vv)], -1)
    ground = torch.stack([0.25 + 0 * vv, 0.2 + 0 * vv, 0.15 + 0 * vv], -1)
    img = torch.where((vv < 0.5)[..., None], sky, ground)
    sun = torch.exp(-(((uu - 0.3) * 2) ** 2 + (vv - 0.25) ** 2) / 0.002) * 60.0
    img = img + sun[..., None] * torch.tensor([1.0, 0.9, 0.7])
    img = img * (0.9 + 0.2 * torch.rand(H, W, 1, generator=g))
    return img.contiguous()


# ----------------------------------------------------------------------------- antialias (K3)


def antialias_pairs(rast, v_pos, faces, mvp):
    """Scalar restatement of nvdiffrast's antialias analysis (un-vendored; Laine et al. 2020 sec. 3.4) for ONE view.

    rast [H,W,4] = (u, v, z/w, tri+1).  Returns a list of (dst_pixel, src_pixel, weight): out[dst] += w*(in[src]-in[dst]).
    For each adjacent pixel pair with different triangle ids: take the front triangle (smaller z/w, or the only one);
    among its silhouette edges (boundary, or the neighbouring triangle's opposite vertex lies on the same side of the
    edge in screen space) that cross the segment between the pixel centres, take the nearest crossing s in [0,1]
    measured from the front triangle's pixel; s > 0.5 blends into the other pixel with s-0.5, s < 0.5 into its own
    pixel with 0.5-s."""
    H, W, _ = rast.shape
    V = v_pos.shape[0]
    clip = torch.cat([v_pos, torch.ones(V, 1)], 1) @ mvp.t()
    sx = (clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * W
    sy = (clip[:, 1] / clip[:, 3] * 0.5 + 0.5) * H
    f = faces.long().tolist()
    edge_map = {}
    for fi, (a, b, c) in enumerate(f):
        for (p, q, o) in ((a, b, c), (b, c, a), (c, a, b)):
            edge_map.setdefault((min(p, q), max(p, q)), []).append((fi, o))
    tri = (rast[..., 3].long() - 1).tolist()
    zw = rast[..., 2].tolist()
    sxl, syl = sx.tolist(), sy.tolist()
    out = []
    for y in range(H):
        for x in range(W):
            for (dy, dx) in ((0, 1), (1, 0)):
                y1, x1 = y + dy, x + dx
                if y1 >= H or x1 >= W:
                    continue
                t0, t1 = tri[y][x], tri[y1][x1]
                if t0 == t1:
                    continue
                if t0 < 0:
                    first = False
                elif t1 < 0:
                    first = True
                else:
                    first = zw[y][x] < zw[y1][x1]
                T = t0 if first else t1
                (cy, cx), (oy, ox) = ((y, x), (y1, x1)) if first else ((y1, x1), (y, x))
                sgn = 1.0 if first else -1.0
                ccx, ccy = cx + 0.5, cy + 0.5
                a, b, c = f[T]
                best = None
                for (p, q, o3) in ((a, b, c), (b, c, a), (c, a, b)):
                    others = [o for (fi, o) in edge_map[(min(p, q), max(p, q))] if fi != T]
                    ex, ey = sxl[q] - sxl[p], syl[q] - syl[p]
                    side_c = ex * (syl[o3] - syl[p]) - ey * (sxl[o3] - sxl[p])
                    if others:
                        oo = others[0]
                        side_o = ex * (syl[oo] - syl[p]) - ey * (sxl[oo] - sxl[p])
                        sil = side_o * side_c > 0
                    else:
                        sil = True
                    if not sil:
                        continue
                    if dx == 1:
                        da, db = syl[p] - ccy, syl[q] - ccy
                        if not da * db < 0:
                            continue
                        s = (sxl[p] + ex * (da / (da - db)) - ccx) * sgn
                    else:
                        da, db = sxl[p] - ccx, sxl[q] - ccx
                        if not da * db < 0:
                            continue
                        s = (syl[p] + ey * (da / (da - db)) - ccy) * sgn
                    if 0 <= s <= 1 and (best is None or s < best):
                        best = s
                if best is None or best == 0.5:
                    continue
                ci, oi = cy * W + cx, oy * W + ox
                out.append((oi, ci, best - 0.5) if best > 0.5 else (ci, oi, 0.5 - best))
    return out


def antialias_apply(x, pairs):
    """x [n_pix, c] -> blended copy (differentiable)."""
    if not pairs:
        return x.clone()
    dst = torch.tensor([p[0] for p in pairs]); src = torch.tensor([p[1] for p in pairs])
    w = torch.tensor([p[2] for p in pairs], dtype=x.dtype)[:, None]
    return x.index_add(0, dst, w * (x[src] - x[dst]))
