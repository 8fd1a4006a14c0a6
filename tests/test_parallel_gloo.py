"""world_size-2 gloo test of the view-parallel step's host logic (no GPU): sharding, the global
normalisers and the single gradient all-reduce must reproduce the single-process gradient.  The per-view
compute is the CPU oracle (render half only, with a quadratic stand-in for the SDS term so that the test
stays in seconds)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from oracle import render as OR
from tests._fixtures import make_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view_grad(sc, b, P, n_grid, meta, env, inv_views, total_pn, target):
    """d/dP of [ mean-over-views 0.5*||rgb - target||^2 + batch-mean mat_reg ] restricted to view b."""
    grid, W1, W2 = P[:n_grid], P[n_grid:n_grid + 2048].view(64, 32), P[n_grid + 2048:].view(5, 64)
    sel = sc["gb"]["selector"][b]
    pts, nrm, vd = sc["gb"]["gb_pos"][b][sel], sc["gb"]["gb_normal"][b][sel], sc["gb"]["gb_viewdirs"][b][sel]
    n = pts.shape[0]
    g = torch.Generator().manual_seed(100 + b)
    ang, eps = torch.rand(n, 1, generator=g), torch.randn(n, 1, generator=g) * 0.05
    rd, rs = torch.rand(n, 1, 1, generator=g), torch.rand(n, 1, 1, generator=g)
    f = OR.geometry_forward(pts, grid, W1, W2, meta)
    fj = OR.geometry_forward(OR.jitter_positions(pts, nrm, ang, eps), grid, W1, W2, meta)
    al, me, ro, _ = OR.material_params(f, fj)
    o = OR.shade_raytracing(pts, nrm, vd, env, me, ro, al, rd, rs, lambda oo, dd: sc["tracer"].trace(oo, dd)[1],
                            n_diffuse=16, n_specular=8)
    m, mj = torch.sigmoid(f), torch.sigmoid(fj)
    kd, ks = (m[:, :3] - mj[:, :3]).abs(), (m[:, 3:5] - mj[:, 3:5]).abs()
    reg = (0.25 * (kd.mean(-1) * kd[:, 2]).sum() + 0.1 * (ks[:, 0] * ks[:, 1]).sum()) / total_pn
    loss = 0.5 * ((o["color"] - target) ** 2).sum() * inv_views + reg
    return loss


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from dreammat_b200.parallel import allreduce_gradients, global_pixel_count, shard_slice
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    V = 4
    sc = make_scene(res=16, subdiv=2, bump=0.1, seed=0, n_views=V)
    meta, total = OR.hashgrid_meta(log2_T=10)   # small table: the test is about the sharding algebra
    n_grid = total * 2
    g = torch.Generator().manual_seed(0)
    P0 = torch.cat([(torch.rand(n_grid, generator=g) * 2 - 1) * 0.3, (torch.rand(2048, generator=g) * 2 - 1) / 5.6,
                    (torch.rand(320, generator=g) * 2 - 1) / 8])
    env = OR.synthetic_envmap(32, 64, 0)
    pn = [int(sc["gb"]["selector"][b].sum()) for b in range(V)]
    view_ids = list(range(V))
    total_pn = global_pixel_count(pn, view_ids)
    mine = view_ids[shard_slice(V, rank, world)]
    P = P0.clone().requires_grad_(True)
    loss = sum(_view_grad(sc, b, P, n_grid, meta, env, 1.0 / V, total_pn, 0.3) for b in mine)
    loss.backward()
    flat = P.grad.clone()
    allreduce_gradients(flat, world)
    if rank == 0:
        Ps = P0.clone().requires_grad_(True)
        full = sum(_view_grad(sc, b, Ps, n_grid, meta, env, 1.0 / V, total_pn, 0.3) for b in view_ids)
        full.backward()
        err = float((flat - Ps.grad).norm() / Ps.grad.norm())
        ret["err"] = err
        ret["nnz"] = int((flat != 0).sum())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_single_process():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["nnz"] > 100
    assert ret["err"] < 1e-5, ret["err"]


def test_shard_slice_validation():
    from dreammat_b200.parallel import shard_slice
    assert shard_slice(8, 3, 4) == slice(6, 8)
    try:
        shard_slice(8, 0, 3)
        assert False
    except ValueError:
        pass


def test_pixel_partition_covers_every_pixel_once():
    from dreammat_b200.parallel import pixel_partition
    for pn, world in (([10, 30, 5, 15], 2), ([10, 30, 5, 15, 7, 9, 100, 3], 4), ([57479, 91134, 197815, 120000, 80000, 64000, 150000, 99000], 8),
                      ([5, 0, 7, 1], 4), ([3, 3], 1)):
        segs, counts = pixel_partition(pn, world)
        per = len(pn) // world
        seen = [[0] * n for n in pn]
        for r in range(world):
            shaded = 0
            for (g, a, b) in segs[r]:
                assert 0 <= a < b <= pn[g]
                for i in range(a, b):
                    seen[g][i] += 1
                shaded += b - a
            assert shaded == sum(counts[r])
            assert abs(shaded - sum(pn) / world) <= 1                      # equal intervals
            # pieces are in global order, so the send buffer is already grouped by owner rank
            owners = [g // per for (g, _, _) in segs[r]]
            assert owners == sorted(owners)
        assert all(c == 1 for row in seen for c in row)
        for o in range(world):   # what rank o receives is exactly its own views' pixels
            assert sum(counts[r][o] for r in range(world)) == sum(pn[o * per:(o + 1) * per])


def _a2a_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from dreammat_b200.parallel import exchange_rows, pixel_partition
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pn = [7, 12, 3, 10]                      # 2 views per rank
    segs, counts = pixel_partition(pn, world)
    off = [0, 7, 19, 22, 32]
    # "colour" of global pixel i is i: every rank shades its interval, owners must end up with their views in order
    send = torch.cat([torch.arange(off[g] + a, off[g] + b, dtype=torch.float32) for (g, a, b) in segs[rank]]).view(-1, 1).repeat(1, 3)
    recv_counts = [counts[r][rank] for r in range(world)]
    own = exchange_rows(send, counts[rank], recv_counts, world)
    lo, hi = off[rank * 2], off[rank * 2 + 2]
    ok = torch.equal(own[:, 0], torch.arange(lo, hi, dtype=torch.float32))
    back = exchange_rows(own * 2, recv_counts, counts[rank], world)           # the backward direction
    ok = ok and torch.equal(back, send * 2)
    from dreammat_b200.parallel import warm_exchange
    warm_exchange(1000, world, "cpu")         # size ladder 16, 64, 256, 1000: terminates, same collectives on every rank
    warm_exchange(3, world, "cpu")
    again = exchange_rows(send, counts[rank], recv_counts, world)             # and leaves the exchange usable
    ok = ok and torch.equal(again, own)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_exchange_rows_roundtrip_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_a2a_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _balanced_worker(rank, world, port, ret):
    """Pixel-balanced step with the CPU oracle as the per-pixel compute: every rank shades an equal interval of the global
    pixel line, colours travel to the view owners (all-to-all), d loss / d colour travels back (transposed all-to-all), the
    parameter gradients meet in the all-reduce.  Must equal the single-process gradient."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from dreammat_b200.parallel import allreduce_gradients, exchange_rows, pixel_partition
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    V = 4
    sc = make_scene(res=16, subdiv=2, bump=0.1, seed=0, n_views=V)
    meta, total = OR.hashgrid_meta(log2_T=10)
    n_grid = total * 2
    g = torch.Generator().manual_seed(0)
    P0 = torch.cat([(torch.rand(n_grid, generator=g) * 2 - 1) * 0.3, (torch.rand(2048, generator=g) * 2 - 1) / 5.6,
                    (torch.rand(320, generator=g) * 2 - 1) / 8])
    env = OR.synthetic_envmap(32, 64, 0)
    pn = [int(sc["gb"]["selector"][b].sum()) for b in range(V)]
    total_pn = sum(pn)
    per = V // world
    segs, counts = pixel_partition(pn, world)
    recv_counts = [counts[r][rank] for r in range(world)]

    def view_arrays(b):
        sel = sc["gb"]["selector"][b]
        gg = torch.Generator().manual_seed(100 + b)          # per-pixel draws are a function of the view only
        n = pn[b]
        return (sc["gb"]["gb_pos"][b][sel], sc["gb"]["gb_normal"][b][sel], sc["gb"]["gb_viewdirs"][b][sel], torch.rand(n, 1, generator=gg),
                torch.randn(n, 1, generator=gg) * 0.05, torch.rand(n, 1, 1, generator=gg), torch.rand(n, 1, 1, generator=gg))

    def shade(P, b, a, bb):
        grid, W1, W2 = P[:n_grid], P[n_grid:n_grid + 2048].view(64, 32), P[n_grid + 2048:].view(5, 64)
        pts, nrm, vd, ang, eps, rd, rs = (t[a:bb] for t in view_arrays(b))
        f = OR.geometry_forward(pts, grid, W1, W2, meta)
        fj = OR.geometry_forward(OR.jitter_positions(pts, nrm, ang, eps), grid, W1, W2, meta)
        al, me, ro, _ = OR.material_params(f, fj)
        o = OR.shade_raytracing(pts, nrm, vd, env, me, ro, al, rd, rs, lambda oo, dd: sc["tracer"].trace(oo, dd)[1], n_diffuse=16, n_specular=8)
        m, mj = torch.sigmoid(f), torch.sigmoid(fj)
        kd, ks = (m[:, :3] - mj[:, :3]).abs(), (m[:, 3:5] - mj[:, 3:5]).abs()
        reg = (0.25 * (kd.mean(-1) * kd[:, 2]).sum() + 0.1 * (ks[:, 0] * ks[:, 1]).sum()) / total_pn
        return o["color"], reg

    def view_loss(color):          # stands in for the SDS term: any function of a view's whole image
        return 0.5 * ((color - 0.3) ** 2).sum() * (1.0 / V) + 0.01 * color.mean() * color.std()

    # ---- balanced, distributed
    P = P0.clone().requires_grad_(True)
    cols, regs = zip(*[shade(P, gv, a, bb) for (gv, a, bb) in segs[rank]])
    color_sh = torch.cat(cols)
    own = exchange_rows(color_sh.detach(), counts[rank], recv_counts, world).requires_grad_(True)
    off, loss_own = 0, 0.0
    for b in range(rank * per, (rank + 1) * per):
        loss_own = loss_own + view_loss(own[off:off + pn[b]]); off += pn[b]
    assert off == own.shape[0]
    (dcol_own,) = torch.autograd.grad(loss_own, own)
    dcol_sh = exchange_rows(dcol_own, recv_counts, counts[rank], world)
    ((color_sh * dcol_sh).sum() + sum(regs)).backward()
    flat = P.grad.clone()
    allreduce_gradients(flat, world)
    if rank == 0:
        Ps = P0.clone().requires_grad_(True)
        full = 0.0
        for b in range(V):
            c, r_ = shade(Ps, b, 0, pn[b])
            full = full + view_loss(c) + r_
        full.backward()
        ret["err"] = float((flat - Ps.grad).norm() / Ps.grad.norm())
        ret["imbalance"] = max(sum(bb - a for (_, a, bb) in s) for s in segs) - min(sum(bb - a for (_, a, bb) in s) for s in segs)
    dist.barrier()
    dist.destroy_process_group()


def test_pixel_balanced_gradient_equals_single_process():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_balanced_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["imbalance"] <= 1
    assert ret["err"] < 1e-5, ret["err"]
