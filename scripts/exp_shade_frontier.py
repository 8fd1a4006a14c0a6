"""A/B of the MC shader's traversal variants on 512x512 views of the 100k-face bench mesh (CUDA events, 1 GPU):
root traversal vs shared-origin frontier traversal, one CTA per 8 pixels vs persistent warps.  Also checks that the
variants produce bit-identical colours (the any-hit result does not depend on the traversal order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreammat_b200 import render_ops as R
from dreammat_b200._cabi import lib
from dreammat_b200.scene import DataConfig, FixCameraSet, procedural_mesh, synthetic_envmap
from dreammat_b200.system import DreamMatMaterial, DreamMatMesh, RaytraceRender
dev = "cuda"
mesh = procedural_mesh(100000, 0.8, 0)
geo = DreamMatMesh({"shape_init": "p"}, dev, mesh=mesh)
mat = DreamMatMaterial({"diffuse_sample_num": 200, "specular_sample_num": 128}, dev, [synthetic_envmap(2048, 4096, 0)])
ren = RaytraceRender({}, geo, mat, None, dev)
cams = FixCameraSet(DataConfig(width=512, height=512), torch.Generator().manual_seed(0))
for vid in (3, 40):
    c = cams.cameras(torch.tensor([vid]))
    g = ren.gbuffer(c["rays_o"].to(dev), c["rays_d"].to(dev), c["mvp_mtx"].to(dev), c["w2c"].to(dev), vid)
    n = g["pn"]
    gen = torch.Generator(device=dev).manual_seed(vid)
    f = torch.randn(n, 5, device=dev, generator=gen); fj = torch.randn(n, 5, device=dev, generator=gen)
    rd, rs = torch.rand(n, device=dev, generator=gen), torch.rand(n, device=dev, generator=gen)
    ref = None
    bvh = R.Bvh(mesh[0], mesh[1])
    for (fr, warps, df, occ, rf) in ((0, 8, 0, 24, 0), (1, 8, 0, 24, 0), (1, 8, 0, 32, 0), (1, 8, 0, 24, 8), (1, 8, 0, 24, 16), (1, 8, 0, 24, 20), (1, 8, 0, 24, 24),
                                     (1, 8, 0, 24, 28), (1, 8, 0, 24, 32), (1, 8, 0, 32, 16), (1, 8, 0, 32, 24), (1, 8, 0, 32, 32)):
        lib().dm_tune(b"mc_refill", rf)
        lib().dm_tune(b"mc_frontier", fr); lib().dm_tune(b"mc_warps", warps); lib().dm_tune(b"mc_defer", df); lib().dm_tune(b"mc_occupancy", occ)
        def run():
            return R.shade_mc(f, fj, g["pts"], g["nrm"], g["vd"], rd, rs, mat.mc_cfg, bvh, mat.light[0], mat.tab_d,
                              mat.tab_s, want_aux=False)[0]
        for _ in range(2): col = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): col = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None: ref = col.clone()
        print(f"view {vid} pn={n} frontier={fr:2d} warps/CTA={warps} defer={df:2d} occupancy={occ} refill={rf:2d}: {ms:.3f} ms  {n*328/ms/1e6:.2f} Grays/s  max|dcol| vs first {float((col - ref).abs().max()):.1e}")
lib().dm_tune(b"mc_frontier", 1); lib().dm_tune(b"mc_warps", 8); lib().dm_tune(b"mc_defer", 0); lib().dm_tune(b"mc_occupancy", 32); lib().dm_tune(b"mc_refill", 0)
