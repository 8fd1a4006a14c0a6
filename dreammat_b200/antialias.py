"""Silhouette antialiasing (nvdiffrast `dr.antialias`; utils/rasterize.py:56 <- models/renderers/
raytracing_renderer.py:127,147,199) as a cached sparse blend.

nvdiffrast is an un-vendored dependency; this restates its published algorithm (Laine et al. 2020, sec. 3.4):
for every horizontally / vertically adjacent pixel pair whose triangle ids differ, the triangle that is in
front (smaller z/w; or the only one) is examined; if one of its edges is a silhouette edge (no neighbour
across it, or the neighbour folds to the same side in screen space) and crosses the segment between the two
pixel centres at parameter s in [0,1] (measured from the front triangle's pixel), the surface is taken to end
there: s > 0.5 bleeds its colour into the other pixel with weight s - 0.5, s < 0.5 lets the other pixel's
colour bleed into its pixel with weight 0.5 - s.

Because the mesh and the training cameras are fixed, the (dst, src, weight) list of a view never changes:
it is built once on the host (vectorised numpy, init-time like the BVH build) and the per-iteration work is
`dm_antialias_fwd` / `dm_antialias_bwd` (out[dst] += w * (in[src] - in[dst]) and its adjoint).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr


def edge_neighbours(faces: np.ndarray, n_verts: int) -> np.ndarray:
    """[F,3] -> [F,3]: for edge e of face f (vertices e, e+1), the vertex opposite to it in the adjacent face
    (-1 for a boundary edge; the first other face for non-manifold edges)."""
    F = faces.shape[0]
    a = faces[:, [0, 1, 2]].reshape(-1).astype(np.int64)
    b = faces[:, [1, 2, 0]].reshape(-1).astype(np.int64)
    opp = faces[:, [2, 0, 1]].reshape(-1).astype(np.int64)
    key = np.minimum(a, b) * n_verts + np.maximum(a, b)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same_next = np.zeros(3 * F, bool)
    same_next[:-1] = ks[:-1] == ks[1:]
    same_prev = np.zeros(3 * F, bool)
    same_prev[1:] = same_next[:-1]
    partner = np.full(3 * F, -1, np.int64)
    idx = np.arange(3 * F)
    partner[same_next] = idx[same_next] + 1
    m = same_prev & ~same_next
    partner[m] = idx[m] - 1
    res_sorted = np.where(partner >= 0, opp[order][np.clip(partner, 0, None)], -1)
    res = np.empty(3 * F, np.int64)
    res[order] = res_sorted
    return res.reshape(F, 3)


def build_pairs(rast: np.ndarray, v_pos: np.ndarray, faces: np.ndarray, nbr_opp: np.ndarray, mvp: np.ndarray
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """rast [H,W,4] = (u, v, z/w, tri+1) -> (dst, src, alpha) with pixel indices y*W + x."""
    H, W, _ = rast.shape
    tri = rast[..., 3].astype(np.int64) - 1
    zw = rast[..., 2]
    clip = np.concatenate([v_pos, np.ones((v_pos.shape[0], 1), v_pos.dtype)], 1) @ mvp.T
    wv = clip[:, 3:4]
    scr = np.stack([(clip[:, 0] / wv[:, 0] * 0.5 + 0.5) * W, (clip[:, 1] / wv[:, 0] * 0.5 + 0.5) * H], -1)
    out = []
    for d in (0, 1):  # 0: horizontal neighbour (x+1), 1: vertical neighbour (y+1)
        if d == 0:
            diff = tri[:, :-1] != tri[:, 1:]
            y0, x0 = np.nonzero(diff)
            y1, x1 = y0, x0 + 1
        else:
            diff = tri[:-1, :] != tri[1:, :]
            y0, x0 = np.nonzero(diff)
            y1, x1 = y0 + 1, x0
        if y0.size == 0:
            continue
        t0, t1 = tri[y0, x0], tri[y1, x1]
        z0, z1 = zw[y0, x0], zw[y1, x1]
        first = np.where(t0 < 0, False, np.where(t1 < 0, True, z0 < z1))      # True: the front triangle is pixel 0's
        T = np.where(first, t0, t1)
        cy, cx = np.where(first, y0, y1), np.where(first, x0, x1)            # pixel of the front triangle
        oy, ox = np.where(first, y1, y0), np.where(first, x1, x0)            # the other pixel
        ccx, ccy = cx + 0.5, cy + 0.5
        sgn = np.where(first, 1.0, -1.0)                                     # direction from c to o along the pair axis
        vi = faces[T]                                                         # [K,3]
        S = scr[vi]                                                           # [K,3,2]
        best = np.full(T.shape[0], np.inf)
        for e in range(3):
            A, B, Cc = S[:, e], S[:, (e + 1) % 3], S[:, (e + 2) % 3]
            opp = nbr_opp[T, e]
            So = scr[np.clip(opp, 0, None)]
            ex, ey = B[:, 0] - A[:, 0], B[:, 1] - A[:, 1]
            side_c = ex * (Cc[:, 1] - A[:, 1]) - ey * (Cc[:, 0] - A[:, 0])
            side_o = ex * (So[:, 1] - A[:, 1]) - ey * (So[:, 0] - A[:, 0])
            sil = (opp < 0) | (side_o * side_c > 0)
            if d == 0:
                da, db = A[:, 1] - ccy, B[:, 1] - ccy
                with np.errstate(divide="ignore", invalid="ignore"):
                    xs = A[:, 0] + ex * (da / (da - db))
                s = (xs - ccx) * sgn
            else:
                da, db = A[:, 0] - ccx, B[:, 0] - ccx
                with np.errstate(divide="ignore", invalid="ignore"):
                    ys = A[:, 1] + ey * (da / (da - db))
                s = (ys - ccy) * sgn
            ok = sil & (da * db < 0) & (s >= 0) & (s <= 1)
            best = np.where(ok & (s < best), s, best)
        ok = np.isfinite(best) & (best != 0.5)
        s = best[ok]
        cidx, oidx = (cy * W + cx)[ok], (oy * W + ox)[ok]
        to_other = s > 0.5
        out.append((np.where(to_other, oidx, cidx), np.where(to_other, cidx, oidx), np.abs(s - 0.5)))
    if not out:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.float32)
    dst = np.concatenate([o[0] for o in out]).astype(np.int32)
    src = np.concatenate([o[1] for o in out]).astype(np.int32)
    alpha = np.concatenate([o[2] for o in out]).astype(np.float32)
    return dst, src, alpha


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dst, src, alpha):
        x = x.float().contiguous()
        c = x.shape[-1]
        n_pix = x.numel() // c
        out = torch.empty_like(x)
        check(lib().dm_antialias_fwd(ptr(x), ptr(dst), ptr(src), ptr(alpha), dst.shape[0], n_pix, c, ptr(out), stream_ptr()),
              "dm_antialias_fwd")
        ctx.save_for_backward(dst, src, alpha)
        return out

    @staticmethod
    def backward(ctx, dout):
        dst, src, alpha = ctx.saved_tensors
        dout = dout.float().contiguous()
        c = dout.shape[-1]
        din = torch.empty_like(dout)
        check(lib().dm_antialias_bwd(ptr(dout), ptr(dst), ptr(src), ptr(alpha), dst.shape[0], dout.numel() // c, c, ptr(din),
                                     stream_ptr()), "dm_antialias_bwd")
        return din, None, None, None


def antialias(x: torch.Tensor, pairs) -> torch.Tensor:
    """x [..., n_pix, c] of ONE view (any leading shape that flattens to pixels), pairs = (dst, src, alpha) tensors."""
    return _Antialias.apply(x, *pairs)
