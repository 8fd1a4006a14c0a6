#!/usr/bin/env python
"""bench.py -- SDS iterations/sec at 512x512 with an 8-view batch (BASELINE.json metric).

One "step" = one full score-distillation iteration over a batch of synthetic views:
PBR Monte-Carlo render (200+128 rays/pixel, BVH occlusion) -> VAE encode (with grad) -> ControlNet + UNet
for the 3 CFG branches -> CSD gradient -> backward through the VAE and the shader into the hash grid /
MLP -> (all-reduce when sharded) -> Adam.  SD-2.1-base / ControlNet / VAE topology with seeded random
weights (no checkpoints offline), fp16 storage + fp32 accumulation like the reference default.

    python bench.py --gpus N --steps K --warmup W            our arm (torchrun for N > 1)
    python bench.py --impl reference ...                     the reference algorithm (CPU oracle port) on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DENSE_TFLOP_PER_VIEW = 5.50    # UNet 2.41 + ControlNet 0.86 + VAE fwd 1.12 + VAE dgrad 1.12
UNET_CN_TFLOP_PER_VIEW = 3.27


def burst_tflops():
    """burst bf16/fp16 tensor peak for a kernel timed alone (MEASURED_PEAKS.json `bf16_tflops`), else the recipe's fallback"""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["bf16_tflops"], "measured"
    except Exception:
        return 1650.0, "fallback"


def kernel_roofline(dtype):
    """Roofline of the dominant tensor kernel, timed ALONE with CUDA events on the launching stream: the CTA-pair
    implicit-GEMM convolution on the UNet's 16x16-latent layer of the 8-view batch (conv3x3 24x16x16, 1280 -> 1280).
    Algorithmic flops per launch = 2 * 6144 * 1280 * 11520; `traffic` is the DRAM read+write of the same launch from the
    committed ncu --set full capture (profiles/r01c_launches_summary.md: 45.3 + 1.3 MB; algorithmic 29.5 MB weights +
    15.7 MB input + 15.7 MB output, the nine taps re-read L2 only)."""
    import torch
    from dreammat_b200 import dense_ops as D
    x = torch.randn(24, 16, 16, 1280, device="cuda").to(dtype)
    w = (torch.randn(1280, 9 * 1280, device="cuda") * 0.01).to(dtype)
    for _ in range(5):
        D.conv2d(x, w, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        D.conv2d(x, w, 3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * 24 * 256 * 1280 * 1280 * 9
    peak, src = burst_tflops()
    ach = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": 46.6e6,
            "kernel": "tc_gemm_pair_kernel<256> conv3x3 24x16x16 1280->1280 (timed alone, %d launches, %.1f us each)" % (n, ms * 1e3),
            "peak_source": src + " burst bf16", "traffic_source": "ncu --set full, profiles/r01c_launches_summary.md"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    def __init__(self, dev):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(dev), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], 0, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ our arm


def build_system(device, res, n_faces, env_hw, seed, dtype):
    import torch
    from dreammat_b200 import weights as Wt
    from dreammat_b200.guidance import PromptProcessorOutput, StableDiffusionLightGuidance
    from dreammat_b200.scene import DataConfig, FixCameraSet, procedural_mesh, synthetic_envmap
    from dreammat_b200.system import DreamMat, DreamMatMaterial, DreamMatMesh, RaytraceRender
    mesh = procedural_mesh(n_faces, 0.8, seed)
    geo = DreamMatMesh({"shape_init": "procedural", "shape_init_params": 0.8}, device, mesh=mesh, seed=seed)
    envs = [synthetic_envmap(env_hw[0], env_hw[1], seed + i) for i in range(5)]
    mat = DreamMatMaterial({"environment_texture": "synthetic", "environment_scale": 2.0, "use_bump": False,
                            "use_raytracing": True, "diffuse_sample_num": 200, "specular_sample_num": 128}, device, envs)
    ren = RaytraceRender({"context_type": "cuda"}, geo, mat, None, device)
    ucfg, vcfg = Wt.UNetConfig(), Wt.VAEConfig()
    gcfg = dict(use_controlnet=True, control_types=["light"], cond_scale=1.05, uncond_scale=[0, -1.0, -0.5, 2000],
                null_scale=[0, 0.0, -0.5, 2000], noise_scale=0.0, min_step_percent=[500, 0.2, 0.02, 501],
                max_step_percent=[500, 0.8, 0.5, 501], control_anneal_start_step=700, condition_scales=[1.0],
                condition_scales_anneal=[0.8])     # configs/dreammat.yaml:54-71
    wu, wc, wv = Wt.random_unet(ucfg, device, 10), Wt.random_controlnet(ucfg, device, 11), Wt.random_vae(vcfg, device, 12)
    guid = StableDiffusionLightGuidance(gcfg, ucfg, vcfg, wu, wc, wv, device, dtype)
    del wu, wc, wv
    torch.cuda.empty_cache()
    g = torch.Generator().manual_seed(seed + 100)
    D = ucfg.cross_attention_dim
    vd, uvd, null = torch.randn(4, 77, D, generator=g), torch.randn(4, 77, D, generator=g), torch.randn(1, 77, D, generator=g)
    pu = PromptProcessorOutput(vd[:1].to(device), uvd[:1].to(device), null.to(device), vd.to(device), uvd.to(device))
    sysm = DreamMat(None, geo, mat, ren, guid, pu, device)
    cams = FixCameraSet(DataConfig(batch_size=1, width=res, height=res), torch.Generator().manual_seed(seed))
    return sysm, cams


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        # NCCL_DEBUG is left exactly as the launcher set it (its banner goes to stderr with everything else, see main())
        dist.init_process_group("nccl", device_id=torch.device(device))
    from dreammat_b200 import _cabi
    _cabi.check(_cabi.lib().dm_device_check(local), "dm_device_check")   # fails loudly without the sm_100a library
    if args.no_pdl:
        _cabi.lib().dm_tune(b"pdl", 0)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    V = args.views
    assert V % world == 0, "global view batch must divide over the ranks"
    Vl = V // world
    sysm, cams = build_system(device, args.res, args.faces, (args.env_h, args.env_w), 0, dtype)
    sysm.world_size, sysm.rank = world, rank
    sysm.balance_pixels = not args.no_balance
    if not args.no_graphs:
        gres = 512 if sysm.resize_to_vae else args.res          # renders that are not 512^2 are resized before the VAE
        sysm.guidance.enable_graphs(Vl, gres, gres)              # dense section as three captured CUDA graphs
    res = args.res
    # a1: per-view camera tensors (fixed set) resident on the device; G-buffers produced once per fixed view
    all_ids = torch.arange(cams.cfg.fix_view_num)
    cam_dev = []
    for v0 in range(0, cams.cfg.fix_view_num, 16):
        c = cams.cameras(all_ids[v0:v0 + 16])
        for j in range(c["mvp_mtx"].shape[0]):
            one = {k: (val[j:j + 1].to(device) if torch.is_tensor(val) else val) for k, val in c.items()}
            sysm.renderer.gbuffer(one["rays_o"], one["rays_d"], one["mvp_mtx"], one["w2c"], v0 + j)
            cam_dev.append({k: one[k] for k in ("mvp_mtx", "w2c", "elevation", "azimuth", "camera_distances")})
    pn = [sysm.renderer._cache[i]["pn"] for i in range(cams.cfg.fix_view_num)]
    sysm.prepare_balanced(range(cams.cfg.fix_view_num))          # one MIN all-reduce: all ranks agree on balanced shading
    # condition maps: synthetic pool standing in for the Blender pre-renders (depth1 + normal3 + 6 x RGB light)
    POOL = 16
    gcond = torch.Generator().manual_seed(7)
    cond_host = torch.rand(POOL, res, res, 22, generator=gcond).pin_memory()
    cond_dev = cond_host.to(device)
    cond_stage = torch.empty(Vl, res, res, 22, device=device)   # per-step H2D landing buffer (e2e leg)
    gsel = torch.Generator().manual_seed(1234)   # shared by all ranks -> the global batch is a function of the step

    def make_batch(from_host):
        view_id, env_id = cams.collate(gsel, V)
        tot_pn = int(sum(pn[int(v)] for v in view_id))
        mine = slice(rank * Vl, (rank + 1) * Vl)
        vid, eid = view_id[mine], env_id[mine]
        b = {"view_id": vid, "env_id": eid, "height": res, "width": res, "rays_o": [None] * Vl, "rays_d": [None] * Vl,
             "global_view_id": view_id, "global_env_id": env_id}   # lets the step balance the shading over the ranks
        for k in ("mvp_mtx", "w2c", "elevation", "azimuth", "camera_distances"):
            b[k] = torch.cat([cam_dev[int(v)][k] for v in vid], 0)
        sel = [(int(v) * 5 + int(e)) % POOL for v, e in zip(vid, eid)]
        if from_host:
            for i, s_ in enumerate(sel):       # pinned host -> device, every step (what Lightning does with the batch)
                cond_stage[i].copy_(cond_host[s_], non_blocking=True)
            b["condition_map"] = cond_stage
        else:
            b["condition_map"] = cond_dev[sel]
        return b, tot_pn

    class _Rays:   # the G-buffer cache is keyed by view id; rays are not needed again
        def __getitem__(self, i):
            return None

    def step(from_host=False):
        b, tot_pn = make_batch(from_host)
        b["rays_o"] = b["rays_d"] = _Rays()
        out = sysm.training_step_fused(b, global_views=V, total_pn_global=tot_pn)
        if from_host:
            return float(out["loss"])      # D2H read of the step's result
        return out["loss"]

    def timed(n_warm, n_steps, from_host):
        for _ in range(n_warm):
            step(from_host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gr = sysm.guidance.graphs
        l0 = _cabi.lib().dm_launch_count() + (gr.replayed_launches if gr else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            step(from_host)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        l1 = _cabi.lib().dm_launch_count() + (gr.replayed_launches if gr else 0)
        return float(ms) / n_steps, (l1 - l0) // n_steps

    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, launches = timed(args.warmup, args.steps, False)
    # section split (one extra profiled step, outside the timed region)
    sec = sysm.profile_step(lambda: make_batch(False), V) if hasattr(sysm, "profile_step") else {}
    ms_e2e, _ = timed(max(1, args.warmup // 2), args.steps, True)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf, src = peaks()
    value = 1000.0 / ms_step
    out = {"metric": "SDS iters/sec at 512x512, 8-view batch", "value": value, "unit": "it/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f16" if dtype == torch.float16 else "bf16",
           "data": "synthetic (procedural %d-face mesh, synthetic HDR env maps and condition maps, seeded random SD-2.1-base/ControlNet/VAE weights)" % args.faces,
           "config": {"workload": "north-star: %dx%d, %d-view batch (%d/GPU), 200+128 MC rays/px, 5 env maps, 128 fixed views" % (res, res, V, Vl),
                      "views": V, "resolution": res, "parallelism": "dp%d (views sharded; shading pixels balanced over ranks by 2 small all-to-alls; 1 all-reduce of 50.4 MB grads)" % world,
                      "l2": "working set (2.5 GB weights + activations) exceeds the 126 MB L2 every step"},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": 1000.0 / ms_e2e, "unit": "it/s", "h2d_bytes_per_step": int(Vl * res * res * 22 * 4),
                   "d2h_bytes_per_step": 4}}
    t_dense = sec.get("dense_ms")
    if t_dense:
        ach = DENSE_TFLOP_PER_VIEW * Vl / (t_dense / 1000.0)
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf,
                           "traffic": None, "kernel": "tc_gemm_pair_kernel + tc_gemm_kernel + attention_kernel over the dense section",
                           "peak_source": src + " sustained bf16",
                           "note": "section-level: algorithmic flops of ALL dense kernels / CUDA-event time of the section inside the timed "
                                   "step (GroupNorm, softmax etc. included); the dominant tensor kernel alone is in roofline_kernel, the "
                                   "MC shader (45 % of the step, issue-bound) in roofline_shading"}
        out["sections_ms"] = sec
        try:
            out["roofline_kernel"] = kernel_roofline(dtype)
        except Exception as ex:  # noqa: BLE001 -- the section-level roofline above stands on its own
            out["roofline_kernel"] = {"error": str(ex)[:200]}
        # shading half: NOT HBM-bound (BVH traversal + FP32 ALU); reported so the fraction is computable:
        # algorithmic bytes = 200 B per covered pixel + one 16 B texel per unoccluded sample (upper bound: every sample)
        t_r = sec.get("render_fwd_ms")
        if t_r:
            pn_step = float(sec.get("pn_local", 0))
            byts = pn_step * (200.0 + 16.0 * 328)
            out["roofline_shading"] = {"bound": "hbm", "achieved": byts / (t_r * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                       "frac": byts / (t_r * 1e-3) / 1e9 / hbm, "traffic": None,
                                       "rays_per_s": pn_step * 328 / (t_r * 1e-3), "covered_pixels": pn_step,
                                       "note": "latency/ALU-bound BVH any-hit traversal; bytes are an upper bound (every sample unoccluded)"}
        out["unet_controlnet_ms_per_step"] = sec.get("unet_cn_ms")
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm


def _cpu_sample(state, res):
    """One bounded sample of the reference algorithm on host cores (oracle port):
    (a) ControlNet+UNet forward for ONE CFG sample at 64x64 latents, (b) VAE encode fwd+bwd at 256x256,
    (c) MC shading + hash-grid fwd/bwd on a 64x64 render.  Returns seconds (a, b, c)."""
    import torch
    from oracle import render as OR
    from oracle import sd as OS
    ucfg, vcfg, wu, wc, wv, sc, grid, W1, W2, meta = state
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    with torch.no_grad():
        z = torch.randn(1, 4, 32, 32, generator=g); t = torch.tensor([500]); ctx = torch.randn(1, 77, 1024, generator=g)
        cond = torch.rand(1, 22, 256, 256, generator=g)
        d, m = OS.controlnet_forward(wc, ucfg, z, t, ctx, cond)
        OS.unet_forward(wu, ucfg, z, t, ctx, d, m)
    t1 = time.perf_counter()
    x = torch.rand(1, 3, 128, 128, generator=g, requires_grad=True)
    mom = OS.vae_encode_moments(wv, vcfg, x)
    mom.square().sum().backward()
    t2 = time.perf_counter()
    gp = grid.clone().requires_grad_(True)
    f = OR.geometry_forward(sc["pts"], gp, W1, W2, meta)
    fj = OR.geometry_forward(OR.jitter_positions(sc["pts"], sc["nrm"], sc["rand_ang"], sc["normal_eps"]), gp, W1, W2, meta)
    al, me, ro, reg = OR.material_params(f, fj)
    out = OR.shade_raytracing(sc["pts"], sc["nrm"], sc["vd"], sc["env"], me, ro, al, sc["rand_d"], sc["rand_s"],
                              lambda o, dd: sc["tracer"].trace(o, dd)[1])
    (out["color"].sum() + reg).backward()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


def _cpu_state():
    import torch
    from oracle import render as OR
    from oracle import sd as OS
    from tests._fixtures import make_scene
    ucfg, vcfg = OS.UNetConfig(), OS.VAEConfig()
    wu, wc, wv = OS.random_unet_weights(ucfg, 10), OS.random_controlnet_weights(ucfg, 11), OS.random_vae_weights(vcfg, 12)
    sc = make_scene(res=32, subdiv=4, bump=0.12, seed=0)
    meta, total = OR.hashgrid_meta()
    g = torch.Generator().manual_seed(0)
    grid = (torch.rand(total * 2, generator=g) * 2 - 1) * 1e-4
    W1 = (torch.rand(64, 32, generator=g) * 2 - 1) / 32 ** 0.5
    W2 = (torch.rand(5, 64, generator=g) * 2 - 1) / 8
    return (ucfg, vcfg, wu, wc, wv, sc, grid, W1, W2, meta)


def _cpu_its(ta, tb, tc, views, res):
    """Extrapolate the bounded sample to one full iteration by pixel count: per view 3 CFG samples at (res/8)^2
    latents (sample: 32^2), VAE at res^2 (sample: 128^2), shading at res^2 (sample: 32^2)."""
    per_view = 3 * ta * (res / 256) ** 2 + tb * (res / 128) ** 2 + tc * (res / 32) ** 2
    return 1.0 / (views * per_view)


SAMPLE_DESC = ("1 CFG sample of ControlNet+UNet @32x32 latents, VAE encode fwd+bwd @128x128, MC shading + hash grid fwd/bwd "
               "@32x32 render (full-size SD-2.1-base topology, fp32); scaled by pixel count to 3 samples/view @64x64 latents, "
               "512^2 VAE, 512^2 render, 8 views")


def cpu_baseline(args):
    import torch
    cores = min(os.cpu_count() or 1, 64)   # torch-CPU conv throughput degrades beyond ~64 threads on this path
    torch.set_num_threads(cores)
    st = _cpu_state()
    ta, tb, tc = _cpu_sample(st, args.res)
    return {"value": _cpu_its(ta, tb, tc, args.views, args.res), "unit": "it/s", "cores": cores, "kind": "port",
            "sample": SAMPLE_DESC, "sample_seconds": {"unet_cn_1sample": ta, "vae_256_fwd_bwd": tb, "shade_64": tc}}


def run_reference(args):
    """The reference's algorithm on the host CPUs (the reference itself is CUDA-only and cannot be installed
    offline: pytorch_lightning / diffusers / nvdiffrast / tinycudann / _raytracing are absent), i.e. the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = min(os.cpu_count() or 1, 64)   # torch-CPU conv throughput degrades beyond ~64 threads on this path
    torch.set_num_threads(cores)
    st = _cpu_state()
    for _ in range(min(args.warmup, 1)):
        _cpu_sample(st, args.res)
    steps = max(1, min(args.steps, 3))   # each sample is ~1 min of CPU work: keep the whole arm to a few minutes
    acc = [0.0, 0.0, 0.0]
    for _ in range(steps):
        s = _cpu_sample(st, args.res)
        acc = [a + b for a, b in zip(acc, s)]
    ta, tb, tc = [a / steps for a in acc]
    its = _cpu_its(ta, tb, tc, args.views, args.res)
    out = {"impl": "reference", "metric": "SDS iters/sec at 512x512, 8-view batch", "value": its, "unit": "it/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": min(args.warmup, 1),
           "ms_per_step": 1000.0 / its, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": {"workload": "north-star: %dx%d, %d-view batch, 200+128 MC rays/px (oracle port on the host cores, bounded sample scaled to the full step)" % (args.res, args.res, args.views),
                                           "views": args.views, "resolution": args.res},
           "cpu_baseline": {"value": its, "unit": "it/s", "cores": cores, "kind": "port", "sample": SAMPLE_DESC},
           "e2e": {"value": its, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_RESULT_FD = None


def emit(obj):
    line = json.dumps(obj) + "\n"
    sys.stdout.flush()
    if _RESULT_FD is None:
        sys.stdout.write(line); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--faces", type=int, default=100000)
    ap.add_argument("--env-h", type=int, default=2048)
    ap.add_argument("--env-w", type=int, default=4096)
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch of the dense kernels")
    ap.add_argument("--no-balance", action="store_true", help="multi-GPU: every rank shades only its own views")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL's version banner, nvcc/ninja chatter)
    # are sent to stderr for the whole run and the result line is written to the saved descriptor.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
