"""View-parallel data parallelism for the SDS step.

Views are independent up to the parameter gradient, so the global batch of V views is split into
contiguous shards of V / world views; SD / ControlNet / VAE weights, env maps, the BVH and the G-buffer
cache are replicated.  The only exchange is ONE all-reduce(SUM) of the flat [hash grid | W1 | W2] gradient
(12 602 288 fp32 = 50.4 MB) per step, after which every rank applies the identical fused Adam update.

Normalisation that keeps the sharded step equal to the single-process one:
  * loss_sds is a mean over the GLOBAL batch of views (dreammat_guidance.py:594 `/ batch_size`), so each
    rank scales its d loss / d latents by 1 / V_global;
  * loss_mat_reg is a mean over the covered pixels of the WHOLE batch (dreammat_material.py:116-117), so
    each rank normalises its pixel sums by the global pixel count.  Every rank holds the full G-buffer
    cache, so that count is known locally -- no extra collective.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_slice(n_views: int, rank: int, world: int) -> slice:
    if n_views % world != 0:
        raise ValueError(f"global view batch {n_views} must divide over {world} ranks")
    per = n_views // world
    return slice(rank * per, (rank + 1) * per)


def global_pixel_count(pn_per_view: Sequence[int], view_ids: Sequence[int]) -> int:
    return int(sum(int(pn_per_view[int(v)]) for v in view_ids))


def allreduce_gradients(flat_grad: torch.Tensor, world: int):
    """The single collective of the step.  NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad
