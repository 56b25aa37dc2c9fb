"""Data-parallel sharding of the inference path (one process per GPU).

Images are independent (eval-mode BN, per-image attention and decode), so a batch shards by
image with no data-path collective; the only exchange is an all-gather of the fixed-shape
per-image detections.  Works with any torch.distributed backend: NCCL over
NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns images [lo, hi); the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local: torch.Tensor, n_images: Optional[int] = None, group=None) -> torch.Tensor:
    """all_gather of ``[b_local, k, 6]`` detections into ``[n_images, k, 6]`` (image order = rank order).
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n_images is None:
        counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
        sizes = [int(c.item()) for c in counts]
    else:
        sizes = [shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0] for r in range(world)]
    assert sizes[rank] == local.shape[0], "local shard does not match the partition"
    m = max(sizes)
    pad = local if local.shape[0] == m else torch.cat(
        (local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))), 0)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


@torch.no_grad()
def detect_sharded(model, images: torch.Tensor, max_det: int = 300, group=None) -> torch.Tensor:
    """``images`` is the GLOBAL batch (same tensor on every rank, any device): each rank runs
    forward + top-k decode on its own slice on its GPU and every rank returns all detections."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(images.shape[0], rank, world)
    dev = model.input_subtract.device
    local = model.detect(images[lo:hi].to(dev, non_blocking=True), max_det=max_det)
    return gather_detections(local, images.shape[0], group)
