"""Shared synthetic scene for the parity tests (seeded; sizes the oracle finishes in seconds)."""
import numpy as np
import torch

from oracle import render as O


def make_scene(res=48, subdiv=3, bump=0.12, seed=0, n_views=1):
    g = torch.Generator().manual_seed(seed)
    v, f = O.icosphere(subdiv, radius=0.8, bump=bump)
    vn = O.vertex_normals(v, f)
    el = torch.tensor([15.0, -10.0, 40.0, 5.0][:n_views])
    az = torch.tensor([30.0, 160.0, -75.0, 100.0][:n_views])
    dist = torch.tensor([3.2, 3.6, 3.9, 3.0][:n_views])
    fovy = torch.tensor([35.0, 30.0, 40.0, 28.0][:n_views])
    cam = O.camera_batch(el, az, dist, fovy, res, res)
    tracer = O.RayTracer(v.numpy(), f.numpy())
    gb = O.gbuffer(tracer, v, f, vn, cam["rays_o"], cam["rays_d"], cam["mvp_mtx"], cam["w2c"])
    sel = gb["selector"]
    pts = gb["gb_pos"][sel]
    nrm = gb["gb_normal"][sel]
    vd = gb["gb_viewdirs"][sel]
    pn = pts.shape[0]
    scene = dict(v=v, f=f, vn=vn, cam=cam, tracer=tracer, gb=gb, pts=pts, nrm=nrm, vd=vd, pn=pn, res=res,
                 features=torch.randn(pn, 5, generator=g), features_jitter=torch.randn(pn, 5, generator=g) * 1.1,
                 rand_d=torch.rand(pn, 1, 1, generator=g), rand_s=torch.rand(pn, 1, 1, generator=g),
                 rand_ang=torch.rand(pn, 1, generator=g), normal_eps=torch.randn(pn, 1, generator=g) * 0.05,
                 env=O.synthetic_envmap(128, 256, seed=seed))
    return scene


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def analytic_envlight(l, roughness=None):
    """Smooth positive stand-in for `envlight.EnvLight.__call__(l, roughness=None)` (third-party, absent): diffuse irradiance
    for `roughness is None`, prefiltered radiance otherwise.  Used by tests/golden/make_splitsum_golden.py when it executes the
    reference's shade_splitsum, and by the test that replays the same inputs through the oracle."""
    a = torch.tensor([[0.9, 0.2, -0.3], [0.1, 0.8, 0.4], [-0.5, 0.3, 0.7]], dtype=l.dtype)
    base = 0.8 + 0.5 * torch.sin(l @ a + torch.tensor([0.3, 1.1, 2.0], dtype=l.dtype))
    if roughness is None:
        return base
    return base * (1.2 - 0.8 * roughness) + 0.15 * roughness * roughness
