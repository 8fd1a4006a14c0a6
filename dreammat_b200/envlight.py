"""Split-sum environment lights, built on the device and cached on disk (rows a5 / N4 of SURVEY.md section 8).

Mirror of `envlight.EnvLight(path, scale)` as the reference constructs it five times at start-up
(models/materials/dreammat_material.py:379-386; ashawkey/envlight @ git HEAD, un-vendored, which wraps nvdiffrec's
renderutils): lat-long HDR -> 128^2 cube -> 2x2-average mips down to 16^2 -> GGX-prefiltered specular mips
(roughness 0.08..0.5 linearly over the chain, 1.0 for the last) + cosine-convolved diffuse cube.  The kernels live in
csrc/envlight.cu; only the NDF-mass cutoff (a one-off 1-D cumulative sum) is computed on the host.
"""
from __future__ import annotations

import hashlib
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr

CUBE_RES = 128      # envlight default max_res
MIN_RES = 16
MIN_ROUGHNESS, MAX_ROUGHNESS = 0.08, 0.5


def ndf_cutoff(roughness: float, cutoff: float = 0.99, n: int = 1000000) -> float:
    """cos of the half-angle inside which the GGX NDF (alpha^2 = roughness^4) holds `cutoff` of its cumulative mass
    (nvdiffrec renderutils.__ndfBounds)."""
    ct = np.cos(np.linspace(0, np.pi / 2.0, n))
    a2 = roughness ** 4
    d = (ct * a2 - ct) * ct + 1.0
    mass = np.cumsum(a2 / (d * d * np.pi))
    return float(ct[int(np.argmax(mass >= mass[-1] * cutoff))])


def latlong_to_cube(latlong: torch.Tensor, res: int, scale: float = 1.0) -> torch.Tensor:
    ll = latlong.float().contiguous()
    cube = torch.empty(6, res, res, 3, device=ll.device)
    check(lib().dm_envlight_latlong_to_cube(ptr(ll), ll.shape[0], ll.shape[1], float(scale), res, ptr(cube), stream_ptr()),
          "dm_envlight_latlong_to_cube")
    return cube


def downsample(cube: torch.Tensor) -> torch.Tensor:
    res = cube.shape[1]
    out = torch.empty(6, res // 2, res // 2, 3, device=cube.device)
    check(lib().dm_envlight_downsample(ptr(cube), res, ptr(out), stream_ptr()), "dm_envlight_downsample")
    return out


def diffuse_cubemap(cube: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(cube)
    check(lib().dm_envlight_filter(ptr(cube), cube.shape[1], 0, 0.0, 0.0, ptr(out), stream_ptr()), "dm_envlight_filter")
    return out


def specular_cubemap(cube: torch.Tensor, roughness: float, cutoff: float = 0.99) -> torch.Tensor:
    out = torch.empty_like(cube)
    check(lib().dm_envlight_filter(ptr(cube), cube.shape[1], 1, float(roughness), ndf_cutoff(roughness, cutoff), ptr(out), stream_ptr()),
          "dm_envlight_filter")
    return out


def build_envlight(latlong_hdr: torch.Tensor, scale: float = 2.0, max_res: int = CUBE_RES, min_res: int = MIN_RES,
                   cache_dir: Optional[str] = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """-> (diffuse cube [6,16,16,3], [specular mips 128, 64, 32, 16]) on the device of `latlong_hdr`.
    cache_dir: results are stored as <md5 of the map bytes and parameters>.pt and re-used (start-up cost, N4)."""
    key = None
    if cache_dir:
        h = hashlib.md5(latlong_hdr.detach().cpu().contiguous().numpy().tobytes())
        h.update(repr((float(scale), max_res, min_res, MIN_ROUGHNESS, MAX_ROUGHNESS, "v1")).encode())
        key = os.path.join(cache_dir, f"envlight_{h.hexdigest()}.pt")
        if os.path.exists(key):
            blob = torch.load(key, map_location=latlong_hdr.device)
            return blob["diffuse"], blob["specular"]
    spec = [latlong_to_cube(latlong_hdr, max_res, scale)]
    while spec[-1].shape[1] > min_res:
        spec.append(downsample(spec[-1]))
    diffuse = diffuse_cubemap(spec[-1])
    n = len(spec)
    for i in range(n - 1):
        r = (i / max(n - 2, 1)) * (MAX_ROUGHNESS - MIN_ROUGHNESS) + MIN_ROUGHNESS
        spec[i] = specular_cubemap(spec[i], r)
    spec[-1] = specular_cubemap(spec[-1], 1.0)
    if key:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = key + f".tmp{os.getpid()}"
        torch.save({"diffuse": diffuse.cpu(), "specular": [m.cpu() for m in spec]}, tmp)
        os.replace(tmp, key)
    return diffuse, spec
