"""GPU parity of the split-sum branch as a PRODUCT path (row a5): device-side envlight build, the reference's real FG LUT
(load/lights/bsdf_256_256.bin) and a real environment map (map1.exr, area-averaged to 64x128 for the fixture) -- both decoded
by the reference's own code in tests/golden/make_splitsum_assets.py -- `use_raytracing=false` through DreamMatMaterial and
through the fused training step.  Reference: models/materials/dreammat_material.py:379-386,405-410,679-711,747-762."""
import os

import pytest
import torch

from oracle import render as O
from tests._fixtures import make_scene, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3
ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "splitsum_assets.pt")


@pytest.fixture(scope="module")
def assets():
    return torch.load(ASSETS)


def test_envlight_device_build_matches_oracle(assets):
    from dreammat_b200 import envlight as E
    ll = assets["envmap_64x128"]
    # reduced chain (32 -> 16 -> 8) so the CPU oracle's O(N^2) prefilter stays in seconds; the kernels are size-generic
    d_o, spec_o = O.build_envlight(ll, scale=2.0, max_res=32, min_res=8)
    d_g, spec_g = E.build_envlight(ll.cuda(), scale=2.0, max_res=32, min_res=8)
    errs = {"diffuse": rel_err(d_g.cpu(), d_o)}
    for i, (a, b) in enumerate(zip(spec_g, spec_o)):
        assert a.shape == b.shape
        errs[f"spec{i}"] = rel_err(a.cpu(), b)
    # base cube alone (no prefilter)
    errs["cube"] = rel_err(E.latlong_to_cube(ll.cuda(), 32, 2.0).cpu(), O.latlong_to_cubemap(ll * 2.0, 32))
    assert abs(E.ndf_cutoff(0.29) - O.ndf_cutoff(0.29)) < 1e-12
    print("\nenvlight device build vs oracle:", " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < 2e-4, errs


def test_envlight_disk_cache_roundtrip(assets, tmp_path):
    from dreammat_b200 import envlight as E
    ll = assets["envmap_64x128"].cuda()
    d1, s1 = E.build_envlight(ll, scale=2.0, max_res=16, min_res=8, cache_dir=str(tmp_path))
    files = os.listdir(tmp_path)
    assert len(files) == 1 and files[0].startswith("envlight_")
    d2, s2 = E.build_envlight(ll, scale=2.0, max_res=16, min_res=8, cache_dir=str(tmp_path))      # served from disk
    assert torch.equal(d1, d2) and all(torch.equal(a, b) for a, b in zip(s1, s2)) and d2.is_cuda
    E.build_envlight(ll, scale=1.0, max_res=16, min_res=8, cache_dir=str(tmp_path))               # other parameters -> other key
    assert len(os.listdir(tmp_path)) == 2


def test_splitsum_product_path_real_lut_and_map(assets):
    """DreamMatMaterial(use_raytracing=false) builds its own lights from the lat-long maps and shades with the real LUT:
    colour, the 7 aux maps, mat_reg and the gradients w.r.t. both feature sets against the oracle (torch autograd)."""
    from dreammat_b200 import envlight as E
    from dreammat_b200.system import DreamMatMaterial
    sc = make_scene(res=48, subdiv=3, bump=0.12, seed=2)
    lut = assets["fg_lut"]
    ll = assets["envmap_64x128"]
    E_CUBE, E_MIN = E.CUBE_RES, E.MIN_RES
    try:
        E.CUBE_RES, E.MIN_RES = 32, 8                      # same reduced chain as the oracle below
        mat = DreamMatMaterial({"use_raytracing": False, "environment_scale": 2.0, "diffuse_sample_num": 200, "specular_sample_num": 128},
                               "cuda", env_maps=[ll, ll * 0.5], fg_lut=lut, envlight=[E.build_envlight((ll * s).cuda(), 2.0, 32, 8) for s in (1.0, 0.5)])
    finally:
        E.CUBE_RES, E.MIN_RES = E_CUBE, E_MIN
    d_o, spec_o = O.build_envlight(ll * 0.5, scale=2.0, max_res=32, min_res=8)
    f = sc["features"].clone().requires_grad_(True)
    fj = sc["features_jitter"].clone().requires_grad_(True)
    albedo, metallic, rough, reg = O.material_params(f, fj, use_raytracing=False)
    out = O.shade_splitsum(sc["nrm"], sc["vd"], d_o, spec_o, lut[0], metallic, rough, albedo)
    g = torch.Generator().manual_seed(5)
    dcol = torch.randn(out["color"].shape, generator=g)
    ((out["color"] * dcol).sum() + 3.0 * reg).backward()
    fc, fjc = sc["features"].cuda().requires_grad_(True), sc["features_jitter"].cuda().requires_grad_(True)
    so, regc = mat(sc["pts"].cuda(), fc, fjc, sc["vd"].cuda(), sc["nrm"].cuda(), torch.tensor(1))
    ((so["color"] * dcol.cuda()).sum() + 3.0 * regc).backward()
    errs = {"color": rel_err(so["color"].detach().cpu(), out["color"].detach()), "reg": abs(float(regc) - float(reg)) / abs(float(reg)),
            "d_features": rel_err(fc.grad.cpu(), f.grad), "d_features_jitter": rel_err(fjc.grad.cpu(), fj.grad)}
    for k in ("albedo", "roughness", "metalness", "specular_lights", "diffuse_lights", "specular_colors", "diffuse_colors"):
        errs[k] = rel_err(so[k].cpu(), out[k].detach())
    print("\nsplit-sum product path (real LUT / map1.exr):", " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < TOL, errs
    with pytest.raises(ValueError):
        DreamMatMaterial({"use_raytracing": False}, "cuda", env_maps=[ll])        # no FG LUT -> loud


def test_fused_step_takes_the_splitsum_branch(assets):
    """`use_raytracing=false` is honoured by training_step_fused (it was ignored): the step's rendered image
    equals the split-sum oracle's, not the Monte-Carlo shader's, and the gradient reaches the hash grid."""
    from dreammat_b200 import envlight as E
    from dreammat_b200 import weights as Wt
    from dreammat_b200.guidance import PromptProcessorOutput, StableDiffusionLightGuidance
    from dreammat_b200.system import DreamMat, DreamMatMaterial, DreamMatMesh, RaytraceRender
    from oracle import sd as OS
    dev, res = "cuda", 64
    sc = make_scene(res=res, subdiv=3, bump=0.12, seed=3, n_views=1)
    lut, ll = assets["fg_lut"], assets["envmap_64x128"]
    g = torch.Generator().manual_seed(4)
    geo = DreamMatMesh({"shape_init": "x"}, dev, mesh=(sc["v"], sc["f"]), seed=5)
    geo.params[:geo.n_grid] = ((torch.rand(geo.n_grid, generator=g) * 2 - 1) * 0.5).to(dev)
    p0 = geo.params.detach().cpu().clone()
    lights = [E.build_envlight(ll.cuda(), 2.0, 32, 8)]
    mat = DreamMatMaterial({"use_raytracing": False, "environment_scale": 2.0}, dev, env_maps=[ll], fg_lut=lut, envlight=lights)
    ren = RaytraceRender({}, geo, mat, None, dev)
    ucfg = OS.UNetConfig(block_out_channels=(64, 128, 128, 128), heads=(1, 2, 2, 2), cross_attention_dim=64)
    vcfg = OS.VAEConfig(block_out_channels=(64, 64, 64, 64))
    guid = StableDiffusionLightGuidance(dict(use_controlnet=True, control_types=["light"], condition_scales=[1.0]),
                                        Wt.UNetConfig(**ucfg.__dict__), Wt.VAEConfig(**vcfg.__dict__), OS.random_unet_weights(ucfg, 0),
                                        OS.random_controlnet_weights(ucfg, 1), OS.random_vae_weights(vcfg, 2), dev)
    vd, uvd, null = torch.randn(4, 77, 64, generator=g), torch.randn(4, 77, 64, generator=g), torch.randn(1, 77, 64, generator=g)
    sysm = DreamMat(None, geo, mat, ren, guid, PromptProcessorOutput(vd[:1], uvd[:1], null, vd, uvd), dev)
    sysm.resize_to_vae = False
    gb = sc["gb"]
    sel = gb["selector"][0]
    pix = torch.nonzero(sel).view(-1).int()
    n = int(pix.shape[0])
    ren._cache[7] = {"pix": pix.to(dev), "pn": n, "pts": gb["gb_pos"][0][sel].to(dev).contiguous(), "nrm": gb["gb_normal"][0][sel].to(dev).contiguous(),
                     "vd": gb["gb_viewdirs"][0][sel].to(dev).contiguous(), "aa": None}
    rng = dict(rand_ang=[torch.rand(n, 1, generator=g)], normal_eps=[torch.randn(n, 1, generator=g) * 0.05], rand_d=[torch.rand(n, 1)],
               rand_s=[torch.rand(n, 1)], t=torch.tensor([400]), noise=torch.randn(1, 4, 8, 8, generator=g), vae_eps=torch.randn(1, 4, 8, 8, generator=g))

    class _N:
        def __getitem__(self, i):
            return None
    batch = {"view_id": torch.tensor([7]), "env_id": torch.tensor([0]), "height": res, "width": res, "rays_o": _N(), "rays_d": _N(),
             "mvp_mtx": sc["cam"]["mvp_mtx"].to(dev), "w2c": sc["cam"]["w2c"].to(dev), "elevation": torch.tensor([15.0]),
             "azimuth": torch.tensor([30.0]), "camera_distances": torch.tensor([3.2]), "condition_map": torch.rand(1, res, res, 22, device=dev)}
    out = sysm.training_step_fused(batch, rng=rng)
    # oracle render of the same view with the split-sum shader
    meta, _ = O.hashgrid_meta()
    grid, W1, W2 = p0[:geo.n_grid], p0[geo.n_grid:geo.n_grid + geo.n_w1].view(64, 32), p0[geo.n_grid + geo.n_w1:].view(5, 64)
    pts, nrm, vdr = gb["gb_pos"][0][sel], gb["gb_normal"][0][sel], gb["gb_viewdirs"][0][sel]
    f = O.geometry_forward(pts, grid, W1, W2, meta)
    fj = O.geometry_forward(O.jitter_positions(pts, nrm, rng["rand_ang"][0], rng["normal_eps"][0]), grid, W1, W2, meta)
    al, me, ro, reg = O.material_params(f, fj, use_raytracing=False)
    d_o, spec_o = [x.cpu() for x in (lights[0][0],)][0], [m.cpu() for m in lights[0][1]]
    col = O.shade_splitsum(nrm, vdr, d_o, spec_o, lut[0], me, ro, al)["color"]
    canvas = torch.ones(res * res, 3).index_put((pix.long(),), col).view(1, res, res, 3)
    e_rgb = rel_err(out["comp_rgb"].cpu(), canvas)
    e_reg = abs(float(out["loss_mat_reg"]) - float(reg)) / abs(float(reg))
    print(f"\nfused step, split-sum branch: rgb {e_rgb:.2e} mat_reg {e_reg:.2e} |grad| {float(geo.grads.abs().sum()):.3e}")
    assert e_rgb < TOL and e_reg < 1e-4
    assert torch.isfinite(geo.grads).all() and float(geo.grads[:geo.n_grid].abs().sum()) > 0 and not torch.equal(geo.params.cpu(), p0)
