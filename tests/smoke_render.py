"""One tiny SDS iteration slice on cuda:0, checked against the CPU oracle (called by __graft_entry__.smoke(); lives under
tests/ because it imports the oracle -- nothing in the dreammat_b200 package does)."""
import torch


def run():
    from oracle import render as O
    from tests._fixtures import make_scene, rel_err
    from dreammat_b200 import render_ops as ops
    from dreammat_b200._cabi import MaterialCfg
    sc = make_scene(res=32, subdiv=2, bump=0.1, seed=1)
    dev = "cuda:0"
    torch.cuda.set_device(0)
    cu = lambda t: t.to(dev).contiguous()
    cfg = ops.default_hashgrid_cfg()
    n_params, _ = ops.hashgrid_num_params(cfg)
    g = torch.Generator().manual_seed(0)
    grid = (torch.rand(n_params, generator=g) * 2 - 1) * 0.5
    W1 = (torch.rand(64, 32, generator=g) * 2 - 1) / 32 ** 0.5
    W2 = (torch.rand(5, 64, generator=g) * 2 - 1)
    meta, _ = O.hashgrid_meta()
    # oracle
    gp = grid.clone().requires_grad_(True)
    pj = O.jitter_positions(sc["pts"], sc["nrm"], sc["rand_ang"], sc["normal_eps"])
    f = O.geometry_forward(sc["pts"], gp, W1, W2, meta)
    fj = O.geometry_forward(pj, gp, W1, W2, meta)
    albedo, metallic, rough, reg = O.material_params(f, fj)
    out = O.shade_raytracing(sc["pts"], sc["nrm"], sc["vd"], sc["env"], metallic, rough, albedo, sc["rand_d"],
                             sc["rand_s"], lambda o, d: sc["tracer"].trace(o, d)[1])
    (out["color"].sum() + reg).backward()
    # device
    gc = cu(grid).requires_grad_(True)
    w1c, w2c = cu(W1), cu(W2)
    pts, nrm = cu(sc["pts"]), cu(sc["nrm"])
    pjc = ops.jitter_positions(pts, nrm, cu(sc["rand_ang"]), cu(sc["normal_eps"]))
    fc = ops.hashgrid_mlp(pts, gc, w1c, w2c, cfg)
    fjc = ops.hashgrid_mlp(pjc, gc, w1c, w2c, cfg)
    mcfg = MaterialCfg(0.0, 0.9, 0.01, 0.9, 200, 128)
    bvh = ops.Bvh(sc["v"], sc["f"])
    color, regc, _ = ops.shade_mc(fc, fjc, pts, nrm, cu(sc["vd"]), cu(sc["rand_d"]), cu(sc["rand_s"]), mcfg, bvh,
                                  ops.envmap_pack(cu(sc["env"])), cu(ops.direction_tables(200)),
                                  cu(ops.direction_tables(128)), want_aux=False)
    (color.sum() + regc).backward()
    torch.cuda.synchronize()
    e1 = rel_err(color.detach().cpu(), out["color"].detach())
    e2 = rel_err(gc.grad.cpu(), gp.grad)
    print(f"smoke: rgb rel err {e1:.2e}, hash-grid grad rel err {e2:.2e}, pn={sc['pn']}")
    assert e1 < 1e-3 and e2 < 5e-3, (e1, e2)
    from tests import smoke_dense
    smoke_dense.run()
