"""Host-side scene plumbing of the hot path: mesh normalisation, the fixed-view camera set and the
per-iteration batch (row a1 of SURVEY.md section 8).  Mirrors, on the host and in torch-CPU ops exactly
like the reference:

    data/uncond.py:584-645,692-698   fixed-view sampling order (elevations, azimuths, distances, fovy)
    data/uncond.py:723-821           `collate`: view_id / env_id draws, c2w / w2c / mvp, rays
    utils/ops.py:179-292             get_ray_directions / get_rays / get_projection_matrix / get_mvp_matrix
    models/geometry/dreammat_mesh.py:142-206  mesh centring, axis alignment, scaling

Blender pre-rendering and mesh file formats stay outside (north_star: untouched).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- meshes


def load_obj(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Minimal Wavefront reader (positions + triangulated faces); trimesh is not on the path."""
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                vs.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, np.float64), np.asarray(fs, np.int64)


def normalize_mesh(v: np.ndarray, scale: float, up: str = "+y", front: str = "+z") -> np.ndarray:
    """dreammat_mesh.py:163-197: centre on the vertex mean, scale max |coord| to `scale`, align to z-up / x-front."""
    dir2vec = {"+x": [1, 0, 0], "+y": [0, 1, 0], "+z": [0, 0, 1], "-x": [-1, 0, 0], "-y": [0, -1, 0], "-z": [0, 0, -1]}
    if up not in dir2vec or front not in dir2vec:
        raise ValueError(f"shape_init_mesh_up and shape_init_mesh_front must be one of {list(dir2vec)}.")
    if up[1] == front[1]:
        raise ValueError("shape_init_mesh_up and shape_init_mesh_front must be orthogonal.")
    v = v - v.mean(0)
    z_, x_ = np.array(dir2vec[up], np.float64), np.array(dir2vec[front], np.float64)
    y_ = np.cross(z_, x_)
    mesh2std = np.linalg.inv(np.stack([x_, y_, z_], axis=0).T)
    v = v / np.abs(v).max() * scale
    return np.dot(mesh2std, v.T).T


def vertex_normals(v: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    i0, i1, i2 = f[:, 0].long(), f[:, 1].long(), f[:, 2].long()
    fn = torch.cross(v[i1] - v[i0], v[i2] - v[i0], dim=-1)
    vn = torch.zeros_like(v)
    vn.index_add_(0, i0, fn); vn.index_add_(0, i1, fn); vn.index_add_(0, i2, fn)
    vn = torch.where((vn * vn).sum(-1, keepdim=True) > 1e-20, vn, torch.tensor([0.0, 0.0, 1.0]))
    return F.normalize(vn, dim=-1)


def procedural_mesh(n_faces_target: int = 100000, scale: float = 0.8, seed: int = 0):
    """Synthetic stand-in for load/shapes/objs/*.obj (absent on the GPU box): a displaced icosphere with
    self-occluding lobes, ~n_faces_target triangles (cat.obj has ~100 k)."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    while f.shape[0] * 4 <= max(n_faces_target, 20) * 1.3:
        # vectorised 1 -> 4 subdivision
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        es = np.sort(e, 1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid], 0)
        nF = f.shape[0]
        ab, bc, ca = base + inv[:nF], base + inv[nF:2 * nF], base + inv[2 * nF:]
        f = np.concatenate([np.stack([f[:, 0], ab, ca], 1), np.stack([f[:, 1], bc, ab], 1),
                            np.stack([f[:, 2], ca, bc], 1), np.stack([ab, bc, ca], 1)], 0)
    rng = np.random.RandomState(seed)
    ph = rng.rand(4) * 6.28
    r = 1 + 0.18 * np.sin(4 * v[:, 0] + ph[0]) * np.cos(3 * v[:, 1] + ph[1]) + 0.12 * np.sin(5 * v[:, 2] + ph[2]) \
        + 0.05 * np.sin(11 * v[:, 0] + 7 * v[:, 1] + ph[3])
    v = v * r[:, None]
    v = v / np.abs(v).max() * scale
    return torch.from_numpy(v.astype(np.float32)), torch.from_numpy(f.astype(np.int32))


# ----------------------------------------------------------------------------- cameras (utils/ops.py)


def get_projection_matrix(fovy: torch.Tensor, aspect_wh: float, near: float = 0.1, far: float = 1000.0):
    """OpenGL-style perspective matrix per view with the y row negated (nvdiffrast's clip space has y down) -- the
    convention of utils/ops.py:266-278; entries [2,2] = -(f+n)/(f-n), [2,3] = -2fn/(f-n), [3,2] = -1."""
    t = torch.tan(fovy.float() / 2.0)
    z, o = torch.zeros_like(t), torch.ones_like(t)
    a22, a23 = -(far + near) / (far - near), -2.0 * far * near / (far - near)
    rows = ((1.0 / (t * aspect_wh), z, z, z), (z, -1.0 / t, z, z), (z, z, o * a22, o * a23), (z, z, -o, z))
    return torch.stack([torch.stack(r, -1) for r in rows], -2)


def get_mvp_matrix(c2w: torch.Tensor, proj: torch.Tensor):
    """world-to-camera as the rigid inverse [R^T | -R^T t] of c2w, and proj @ w2c (utils/ops.py:281-292) -> (mvp, w2c)."""
    Rt = c2w[:, :3, :3].transpose(1, 2)
    top = torch.cat([Rt, -(Rt @ c2w[:, :3, 3:])], -1)
    last = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=c2w.dtype, device=c2w.device).expand(c2w.shape[0], 1, 4)
    w2c = torch.cat([top, last], 1)
    return proj @ w2c, w2c


def get_rays(c2w: torch.Tensor, focal: torch.Tensor, H: int, W: int):
    """utils/ops.py:179-259 with pixel centres at +0.5, normalised directions (uncond.py:780-786)."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32) + 0.5, torch.arange(H, dtype=torch.float32) + 0.5,
                          indexing="xy")
    d = torch.stack([(i - W / 2), -(j - H / 2), -torch.ones_like(i)], -1)[None].repeat(c2w.shape[0], 1, 1, 1)
    d[..., :2] = d[..., :2] / focal[:, None, None, None]
    rays_d = F.normalize((d[..., None, :] * c2w[:, None, None, :3, :3]).sum(-1), dim=-1)
    rays_o = c2w[:, None, None, :3, 3].expand(rays_d.shape)
    return rays_o.contiguous(), rays_d.contiguous()


@dataclass
class DataConfig:
    """configs/dreammat.yaml:6-25 (`random-camera-datamodule`), fields used on the hot path."""
    batch_size: int = 1
    width: int = 512
    height: int = 512
    camera_distance_range: Tuple[float, float] = (3.0, 4.0)
    fovy_range: Tuple[float, float] = (25, 45)
    elevation_range: Tuple[float, float] = (-20, 45)
    azimuth_range: Tuple[float, float] = (-180, 180)
    fix_view_num: int = 128
    fix_env_num: int = 5


class FixCameraSet:
    """The 128 fixed views of FixCameraIterableDataset (uncond.py:584-645, 692-698).  Draw order follows the reference
    (elevations x2, azimuths, distances, camera / centre / up perturbs, fovy) so a shared CPU seed gives the same set --
    pinned by tests/test_oracle_golden.py against the reference's own set_fix_* methods.  The perturbs are 0 in
    dreammat.yaml: their three draws are consumed (they advance the generator before fovy) and discarded."""

    def __init__(self, cfg: DataConfig, generator: Optional[torch.Generator] = None):
        self.cfg = cfg
        g = generator
        n = cfg.fix_view_num
        e0, e1 = cfg.elevation_range
        el1 = torch.rand(n // 2, generator=g) * (e1 - e0) + e0
        pr = [(e0 + 90.0) / 180.0, (e1 + 90.0) / 180.0]
        el2 = torch.asin(2 * (torch.rand(n - n // 2, generator=g) * (pr[1] - pr[0]) + pr[0]) - 1.0) / math.pi * 180.0
        self.elevation_deg = torch.cat((el1, el2))
        a0, a1 = cfg.azimuth_range
        self.azimuth_deg = (torch.rand(n, generator=g) + torch.arange(n)) / n * (a1 - a0) + a0
        d0, d1 = cfg.camera_distance_range
        self.camera_distances = torch.rand(n, generator=g) * (d1 - d0) + d0
        torch.rand(n, 3, generator=g)       # camera_perturbs * 0   (uncond.py:623-628)
        torch.randn(n, 3, generator=g)      # center_perturbs * 0   (:630-633)
        torch.randn(n, 3, generator=g)      # up_perturbs * 0       (:635-639)
        f0, f1 = cfg.fovy_range
        self.fovy_deg = torch.rand(n, generator=g) * (f1 - f0) + f0

    def cameras(self, view_ids: torch.Tensor) -> Dict[str, torch.Tensor]:
        """uncond.py:740-796 for the selected views."""
        cfg = self.cfg
        el_deg, az_deg = self.elevation_deg[view_ids], self.azimuth_deg[view_ids]
        dist, fovy_deg = self.camera_distances[view_ids], self.fovy_deg[view_ids]
        el, az = el_deg * math.pi / 180, az_deg * math.pi / 180
        pos = torch.stack([dist * torch.cos(el) * torch.cos(az), dist * torch.cos(el) * torch.sin(az), dist * torch.sin(el)], -1)
        center = torch.zeros_like(pos)
        up = torch.as_tensor([0, 0, 1], dtype=torch.float32)[None].repeat(pos.shape[0], 1)
        lookat = F.normalize(center - pos, dim=-1)
        right = F.normalize(torch.cross(lookat, up, dim=-1), dim=-1)
        up = F.normalize(torch.cross(right, lookat, dim=-1), dim=-1)
        c2w3x4 = torch.cat([torch.stack([right, up, -lookat], dim=-1), pos[:, :, None]], dim=-1)
        c2w = torch.cat([c2w3x4, torch.zeros_like(c2w3x4[:, :1])], dim=1)
        c2w[:, 3, 3] = 1.0
        fovy = fovy_deg * math.pi / 180
        focal = 0.5 * cfg.height / torch.tan(0.5 * fovy)
        rays_o, rays_d = get_rays(c2w, focal, cfg.height, cfg.width)
        proj = get_projection_matrix(fovy, cfg.width / cfg.height, 0.1, 1000.0)
        mvp, w2c = get_mvp_matrix(c2w, proj)
        return {"rays_o": rays_o, "rays_d": rays_d, "mvp_mtx": mvp, "camera_positions": pos, "c2w": c2w, "w2c": w2c,
                "light_positions": pos, "elevation": el_deg, "azimuth": az_deg, "camera_distances": dist,
                "height": cfg.height, "width": cfg.width}

    def collate(self, generator: Optional[torch.Generator] = None, batch_size: Optional[int] = None):
        """uncond.py:723-725, 797: view_id, env_id ~ floor(rand * n) on the CPU generator."""
        B = batch_size or self.cfg.batch_size
        view_id = torch.floor(torch.rand(B, generator=generator) * self.cfg.fix_view_num).long()
        env_id = torch.floor(torch.rand(B, generator=generator) * self.cfg.fix_env_num).long()
        return view_id, env_id


class FixViewMaps:
    """The pre-rendered condition maps of `FixCameraIterableDataset.render_fixview_imgs` (data/uncond.py:532-582), read
    from the reference's directory layout

        <root>/depth/%03d.png                         16-bit PNG, millimetres
        <root>/normal/%03d.png                        8-bit RGB
        <root>/light/%03d_m{0.0,1.0}r{0.0,0.5,1.0}_env{1..E}.png   8-bit RGB, six per (view, env)

    with the reference's arithmetic: RGB = INTER_AREA resize -> BGR2RGB -> / 255; depth = / 1000 -> INTER_NEAREST resize ->
    inverse depth of the covered pixels rescaled to [0.3, 1] by their min / max (:540-557).  Row N1 of SURVEY.md section 8f:
    the RGB maps are kept as the uint8 the resize produced (3.0 GB for 128 views x 5 envs at 512^2 instead of 12.1 GB of
    fp32 -- `x / 255` of the stored byte is bit-identical to the reference's float map), optionally resident on the device,
    and `condition_map` gathers [depth | normal | m0r0 m0r.5 m0r1 m1r0 m1r.5 m1r1] for the (view, env) pairs of a batch
    (:799-802)."""

    LIGHT_TAGS = ("m0.0r0.0", "m0.0r0.5", "m0.0r1.0", "m1.0r0.0", "m1.0r0.5", "m1.0r1.0")

    def __init__(self, root: str, n_views: int, n_envs: int, height: int, width: int, device="cpu"):
        import cv2
        dim = (width, height)
        self.depths = torch.zeros(n_views, height, width, 1)
        self.normals = torch.zeros(n_views, height, width, 3, dtype=torch.uint8)
        self.lightmaps = torch.zeros(n_views, n_envs, height, width, 18, dtype=torch.uint8)

        def rgb_u8(path):
            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise FileNotFoundError(path)
            img = cv2.resize(img, dim, interpolation=cv2.INTER_AREA)
            return torch.from_numpy(np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB)))

        def depth_f32(path):
            raw = cv2.imread(path, cv2.IMREAD_ANYDEPTH)
            if raw is None:
                raise FileNotFoundError(path)
            depth = cv2.resize(raw / 1000, dim, interpolation=cv2.INTER_NEAREST)
            mask = depth > 0
            if mask.sum() <= 0:
                return torch.from_numpy(depth[..., None])
            inv = 1.0 / (depth + 1e-6)
            dmax, dmin = inv[mask].max(), inv[mask].min()
            depth[mask] = (1 - 0.3) * (inv[mask] - dmin) / (dmax - dmin + 1e-6) + 0.3
            return torch.from_numpy(depth[..., None])

        for v in range(n_views):
            self.depths[v] = depth_f32(os.path.join(root, "depth", f"{v:03d}.png"))
            self.normals[v] = rgb_u8(os.path.join(root, "normal", f"{v:03d}.png"))
            for e in range(1, n_envs + 1):
                self.lightmaps[v, e - 1] = torch.cat([rgb_u8(os.path.join(root, "light", f"{v:03d}_{tag}_env{e}.png"))
                                                      for tag in self.LIGHT_TAGS], -1)
        self.to(device)

    def to(self, device):
        self.depths, self.normals, self.lightmaps = (t.to(device) for t in (self.depths, self.normals, self.lightmaps))
        return self

    @classmethod
    def synthetic(cls, n_views: int, n_envs: int, height: int, width: int, device="cpu", seed: int = 0):
        """Stand-in for the Blender pre-renders where they do not exist (benchmarks): same tensors, dtypes and value
        ranges -- depth fp32 in {0} U [0.3, 1], normal / light uint8 -- generated on `device`."""
        self = cls.__new__(cls)
        g = torch.Generator(device=device).manual_seed(seed)
        d = torch.rand(n_views, height, width, 1, device=device, generator=g)
        self.depths = torch.where(d > 0.35, 0.3 + 0.7 * d, torch.zeros_like(d)).contiguous()
        self.normals = torch.randint(0, 256, (n_views, height, width, 3), device=device, generator=g, dtype=torch.uint8)
        self.lightmaps = torch.randint(0, 256, (n_views, n_envs, height, width, 18), device=device, generator=g, dtype=torch.uint8)
        return self

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.depths, self.normals, self.lightmaps))

    def condition_map(self, view_id: torch.Tensor, env_id: torch.Tensor) -> torch.Tensor:
        """[B, H, W, 22] fp32 on the maps' device; channel order of uncond.py:581-582, :802."""
        v, e = view_id.to(self.depths.device), env_id.to(self.depths.device)
        return torch.cat((self.depths[v], self.normals[v].float() / 255.0, self.lightmaps[v, e].float() / 255.0), -1)


def load_hdr_image(path: str) -> torch.Tensor:
    """`load_hdr_image` (dreammat_material.py:65-68): lat-long radiance map (EXR / HDR) as RGB float32 [H, W, 3]."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    img = cv2.imread(path, cv2.IMREAD_ANYCOLOR | cv2.IMREAD_ANYDEPTH)
    if img is None:
        raise FileNotFoundError(path)
    return torch.from_numpy(np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))).float()


def load_reference_envmaps(environment_texture: str, n: int = 5):
    """The five lat-long maps `DreamMatMaterial.configure` reads (dreammat_material.py:379-386):
    <environment_texture>/map{i}/map{i}.exr, i = 1..5 -> list of [H, W, 3] float32 (the `env_maps` argument of
    dreammat_b200.system.DreamMatMaterial)."""
    return [load_hdr_image(os.path.join(environment_texture, f"map{i}", f"map{i}.exr")) for i in range(1, n + 1)]


def load_fg_lut(path: str = "load/lights/bsdf_256_256.bin") -> torch.Tensor:
    """The split-sum BRDF table `FG_LUT` (dreammat_material.py:405-410): 256 x 256 x 2 float32, [1, 256, 256, 2]."""
    a = np.fromfile(path, dtype=np.float32)
    if a.size != 256 * 256 * 2:
        raise ValueError(f"{path}: expected 131072 float32 values, found {a.size}")
    return torch.from_numpy(a.reshape(1, 256, 256, 2))


def synthetic_envmap(H=512, W=1024, seed=0) -> torch.Tensor:
    """HDR lat-long map standing in for load/lights/envmap/map{1..5}.exr (100 MB each; absent on the GPU box)."""
    g = torch.Generator().manual_seed(seed)
    v = (torch.arange(H, dtype=torch.float32) + 0.5) / H
    u = (torch.arange(W, dtype=torch.float32) + 0.5) / W
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    sky = torch.stack([0.4 + 0.3 * (1 - vv), 0.5 + 0.3 * (1 - vv), 0.7 + 0.5 * (1 - vv)], -1)
    ground = torch.stack([0.25 + 0 * vv, 0.2 + 0 * vv, 0.15 + 0 * vv], -1)
    img = torch.where((vv < 0.5)[..., None], sky, ground)
    su, sv = 0.15 + 0.7 * float(torch.rand(1, generator=g)), 0.15 + 0.2 * float(torch.rand(1, generator=g))
    sun = torch.exp(-(((uu - su) * 2) ** 2 + (vv - sv) ** 2) / 0.002) * 60.0
    img = img + sun[..., None] * torch.tensor([1.0, 0.9, 0.7])
    return (img * (0.9 + 0.2 * torch.rand(H, W, 1, generator=g))).contiguous()
