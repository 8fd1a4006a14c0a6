"""Feasibility: does the SIMT MC shader overlap with the tensor-core dense section when launched on two streams?
Times (a) the three dense graphs alone, (b) the shading of the batch alone, (c) both at once (CUDA events around the join)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from dreammat_b200._cabi import check, lib, ptr, stream_ptr
dev = "cuda"
V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sysm, cams = bench.build_system(dev, 512, 100000, (2048, 4096), 0, torch.float16, "mc")
guid, mat, ren, geo = sysm.guidance, sysm.material, sysm.renderer, sysm.geometry
g = guid.enable_graphs(V, 512, 512)
views = list(range(V))
for v in views:
    c = cams.cameras(torch.tensor([v]))
    ren.gbuffer(c["rays_o"].to(dev), c["rays_d"].to(dev), c["mvp_mtx"].to(dev), c["w2c"].to(dev), v)
bufs = []
for v in views:
    ge = ren._cache[v]; n = ge["pn"]
    bufs.append((ge, n, torch.randn(n, 5, device=dev), torch.randn(n, 5, device=dev), torch.rand(n, device=dev), torch.rand(n, device=dev),
                 torch.empty(n, 3, device=dev), torch.empty(n, 9, device=dev)))
reg = torch.zeros(2, device=dev)
env = mat.light[0]

def shade_all():
    st = stream_ptr()
    for (ge, n, f, fj, rd, rs, col, jac) in bufs:
        check(lib().dm_shade_mc_fwd(C.byref(mat.mc_cfg), ren.ray_tracer.h, ptr(env), env.shape[0], env.shape[1], ptr(mat.tab_d), ptr(mat.tab_s),
                                    ptr(ge["pts"]), ptr(ge["nrm"]), ptr(ge["vd"]), ptr(f), ptr(fj), ptr(rd), ptr(rs), n, ptr(col), ptr(jac), ptr(reg),
                                    *([None] * 7), None, None, st), "shade")

def dense_all():
    g.g_vae.replay(); g.g_unet.replay(); g.g_bwd.replay()

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

s2 = torch.cuda.Stream()
def both():
    cur = torch.cuda.current_stream()
    s2.wait_stream(cur)
    with torch.cuda.stream(s2):
        shade_all()
    dense_all()
    cur.wait_stream(s2)

ta, tb = timeit(dense_all), timeit(shade_all)
tc = timeit(both)
print(f"views={V}: dense alone {ta:.2f} ms, shading alone {tb:.2f} ms, sum {ta+tb:.2f} ms, concurrent on two streams {tc:.2f} ms "
      f"(hidden: {ta+tb-tc:.2f} ms = {100*(ta+tb-tc)/min(ta,tb):.0f} % of the shorter one)")
for occ in (24,):
    lib().dm_tune(b"mc_occupancy", occ)
    tc2 = timeit(both)
    print(f"  mc_occupancy={occ}: concurrent {tc2:.2f} ms")
lib().dm_tune(b"mc_occupancy", 32)
for cps in (1,):
    lib().dm_tune_gemm(cps)
    g2 = guid.enable_graphs(V, 512, 512)
    g = g2
    ta2 = timeit(dense_all); tc3 = timeit(both)
    print(f"  gemm CTAs/SM knob {cps}: dense alone {ta2:.2f} ms, concurrent {tc3:.2f} ms")
