"""Data-parallel sharding of the inference path (one process per GPU).

Images are independent (eval-mode BN, per-image attention and decode), so a batch shards by
image with no data-path collective; the only exchange is an all-gather of the fixed-shape
per-image detections.  Works with any torch.distributed backend: NCCL over
NVLink on the GPU box, gloo in the CPU tests.

The gather is latency-bound (1.8 MB per rank at 256 images) and rank-synchronous, so it must
not sit on the compute stream: ``ShardedDetector`` issues ``all_gather_into_tensor`` into a
pre-allocated ring of output buffers on a side stream and hands back a ticket; the next
forward is launched meanwhile and only a consumer of the gathered tensor waits for it.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns images [lo, hi); the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_detections(local: torch.Tensor, n_images: Optional[int] = None, group=None) -> torch.Tensor:
    """all-gather of ``[b_local, k, 6]`` detections into ``[n_images, k, 6]`` (image order = rank order), one
    ``all_gather_into_tensor`` into a single buffer.  Uneven shards are padded to the largest shard for the
    collective and trimmed afterwards.  Synchronous with respect to the current stream (see ``ShardedDetector``
    for the overlapped form)."""
    world, rank = _world(group)
    if world == 1:
        return local
    if n_images is None:
        counts = torch.zeros(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
        sizes = [int(c) for c in counts.tolist()]
    else:
        sizes = [shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0] for r in range(world)]
    if sizes[rank] != local.shape[0]:
        raise ValueError(f"local shard holds {local.shape[0]} images, the partition gives rank {rank} {sizes[rank]}")
    m = max(sizes)
    pad = local if local.shape[0] == m else torch.cat(
        (local, local.new_zeros((m - local.shape[0],) + tuple(local.shape[1:]))), 0)
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if min(sizes) == m:
        return out
    return torch.cat([out[r * m:r * m + s] for r, s in enumerate(sizes)], 0)


class ShardedDetector:
    """forward + decode on this rank's shard, detections of ALL ranks gathered off the compute stream.

    ``submit(local_images)`` runs ``model.detect`` on the calling rank's images (equal shard sizes on every rank)
    and launches the all-gather of the ``[b, k, 6]`` result on a side stream into one of ``depth`` pre-allocated
    ``[world * b, k, 6]`` buffers; ``collect(ticket)`` makes the current stream wait for that gather and returns
    the buffer (valid until ``depth`` further submits).  With depth >= 2 the gather of step i overlaps the forward
    of step i+1, so a slow rank delays the others by at most one step's slack instead of every step.
    """

    def __init__(self, model, max_det: int = 300, group=None, depth: int = 2):
        self.model, self.max_det, self.group, self.depth = model, int(max_det), group, max(1, int(depth))
        self.world, self.rank = _world(group)
        self._bufs: List[torch.Tensor] = []
        self._events: List[Optional[torch.cuda.Event]] = []
        self._locals: List[Optional[torch.Tensor]] = []
        self._side: Optional[torch.cuda.Stream] = None
        self._n = 0

    def _ensure(self, local: torch.Tensor) -> None:
        shape = (self.world * local.shape[0],) + tuple(local.shape[1:])
        if not self._bufs or tuple(self._bufs[0].shape) != shape or self._bufs[0].device != local.device:
            self._bufs = [torch.empty(shape, dtype=local.dtype, device=local.device) for _ in range(self.depth)]
            self._events = [None] * self.depth
            self._locals = [None] * self.depth
            self._side = torch.cuda.Stream(device=local.device) if local.is_cuda else None

    @torch.no_grad()
    def submit(self, local_images: torch.Tensor) -> int:
        local = self.model.detect(local_images, max_det=self.max_det)
        return self.submit_detections(local)

    def submit_detections(self, local: torch.Tensor) -> int:
        """Same, for detections computed by the caller (any fixed-shape ``[b, k, 6]`` tensor, e.g. the NMS output)."""
        ticket = self._n
        self._n += 1
        if self.world == 1:
            self._bufs, self._events = [local], [None]
            return ticket
        self._ensure(local)
        slot = ticket % self.depth
        if local.is_cuda:
            main = torch.cuda.current_stream(local.device)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                dist.all_gather_into_tensor(self._bufs[slot], local.contiguous(), group=self.group)
                done = torch.cuda.Event()
                done.record(self._side)
            local.record_stream(self._side)
            self._events[slot], self._locals[slot] = done, local
        else:   # gloo / CPU tensors: nothing to overlap with
            dist.all_gather_into_tensor(self._bufs[slot], local.contiguous(), group=self.group)
        return ticket

    def collect(self, ticket: int) -> torch.Tensor:
        if self.world == 1:
            return self._bufs[0]
        if ticket < self._n - self.depth:
            raise RuntimeError(f"ticket {ticket} has been overwritten (depth {self.depth})")
        slot = ticket % self.depth
        ev = self._events[slot]
        if ev is not None:
            torch.cuda.current_stream(self._bufs[slot].device).wait_event(ev)
        return self._bufs[slot]

    def __call__(self, local_images: torch.Tensor) -> torch.Tensor:
        return self.collect(self.submit(local_images))


@torch.no_grad()
def detect_sharded(model, images: torch.Tensor, max_det: int = 300, group=None) -> torch.Tensor:
    """``images`` is the GLOBAL batch (same tensor on every rank, any device): each rank runs
    forward + top-k decode on its own slice on its GPU and every rank returns all detections."""
    world, rank = _world(group)
    lo, hi = shard_range(images.shape[0], rank, world)
    dev = model.input_subtract.device
    local = model.detect(images[lo:hi].to(dev, non_blocking=True), max_det=max_det)
    return gather_detections(local, images.shape[0], group)
