"""Multi-GPU plumbing (one process per GPU, torch.distributed).  Rendering shards by image row bands
with no data-path collective; training shards the ray batch and all-reduces the gradients once."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def row_band(rank: int, world: int, height: int) -> Tuple[int, int]:
    """(row0, n_rows) of rank's contiguous band; bands tile [0, height) exactly and differ by <= 1 row."""
    if not (0 <= rank < world) or height < 0:
        raise ValueError("row_band: bad rank/world/height")
    lo, hi = height * rank // world, height * (rank + 1) // world
    return lo, hi - lo


def ray_shard(rank: int, world: int, n_rays: int) -> Tuple[int, int]:
    """(first, count) of rank's slice of a ray batch."""
    return row_band(rank, world, n_rays)


def allreduce_sum_(tensors: Iterable[torch.Tensor], extra: Optional[torch.Tensor] = None,
                   group=None) -> Optional[torch.Tensor]:
    """Sum-all-reduce a list of tensors in place as ONE flat bucket (the 2 x 530,052 fp32 gradients are
    4.24 MB: latency-bound, so one collective).  ``extra`` (e.g. the loss) rides along; returns its
    reduced value.  No-op without an initialised process group or with world size 1."""
    tensors = list(tensors)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return extra
    if len(tensors) == 1 and extra is None and tensors[0].is_contiguous():       # already a bucket: in place
        dist.all_reduce(tensors[0], group=group)
        return None
    parts = [t.reshape(-1) for t in tensors]
    if extra is not None:
        parts.append(extra.reshape(-1).to(parts[0].dtype))
    flat = torch.cat(parts)
    dist.all_reduce(flat, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return flat[off:].reshape(extra.shape) if extra is not None else None


def broadcast_parameters_(tensors: Iterable[torch.Tensor], src: int = 0, group=None) -> None:
    """Make every rank's replicas equal to rank ``src``'s, in place, as ONE flat broadcast (data-parallel training: the
    replicas must start -- and restart after a checkpoint load -- from identical weights, or the summed gradients belong
    to no single model).  No-op without an initialised process group or with world size 1."""
    tensors = [t.data if isinstance(t, torch.nn.Parameter) else t for t in tensors]
    if not tensors or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


def gather_rows(band: torch.Tensor, height: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Optional: assemble the full image on ``dst`` from every rank's band (rows may differ by one, so the
    bands are padded to the largest).  Not on the timed path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return band
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [row_band(r, world, height)[1] for r in range(world)]
    pad = max(sizes)
    buf = band.new_zeros((pad,) + tuple(band.shape[1:]))
    buf[:band.shape[0]] = band
    out: Optional[List[torch.Tensor]] = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
