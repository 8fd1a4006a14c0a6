"""Time dm_shade_mc_fwd on one 512x512 view of the 100k-face bench mesh (perm on/off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreammat_b200 import render_ops as R
from dreammat_b200.scene import DataConfig, FixCameraSet, procedural_mesh, synthetic_envmap
from dreammat_b200.system import DreamMatMaterial, DreamMatMesh, RaytraceRender
dev = "cuda"
mesh = procedural_mesh(100000, 0.8, 0)
geo = DreamMatMesh({"shape_init": "p"}, dev, mesh=mesh)
mat = DreamMatMaterial({"diffuse_sample_num": 200, "specular_sample_num": 128}, dev, [synthetic_envmap(2048, 4096, 0)])
ren = RaytraceRender({}, geo, mat, None, dev)
cams = FixCameraSet(DataConfig(width=512, height=512), torch.Generator().manual_seed(0))
for vid in (3,):
    c = cams.cameras(torch.tensor([vid]))
    g = ren.gbuffer(c["rays_o"].to(dev), c["rays_d"].to(dev), c["mvp_mtx"].to(dev), c["w2c"].to(dev), vid)
    n = g["pn"]
    f = torch.randn(n, 5, device=dev); fj = torch.randn(n, 5, device=dev)
    rd, rs = torch.rand(n, device=dev), torch.rand(n, device=dev)
    from dreammat_b200._cabi import lib
    for name, leaf in (("leaf<=4", 4), ("leaf<=2", 2), ("leaf<=1", 1), ("leaf<=3", 3)):
        lib().dm_tune(b"bvh_leaf", leaf)
        bvh = R.Bvh(mesh[0], mesh[1])
        pm = None
        def run():
            return R.shade_mc(f, fj, g["pts"], g["nrm"], g["vd"], rd, rs, mat.mc_cfg, bvh, mat.light[0], mat.tab_d,
                              mat.tab_s, want_aux=False, perm=pm)[0]
        for _ in range(2): col = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): col = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"view {vid} pn={n} {name:30s} perm={'morton' if pm is not None else 'id'}: {ms:.3f} ms  {n*328/ms/1e6:.2f} Grays/s  checksum {float(col.sum()):.4f}")
