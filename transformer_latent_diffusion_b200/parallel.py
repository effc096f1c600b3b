"""Batch sharding for multi-GPU sampling: one process per GPU, every image's cond+uncond CFG pair
stays on one GPU, **no collective on the data path**; the only communication is an optional gather of the finished
images/latents to one rank.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).

The reference has no multi-GPU inference code (its only parallelism is DDP training, tld/train.py:69,109); this is
the batch-axis sharding BASELINE.json's north_star asks for.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, end) of `rank`'s contiguous shard; the first n_items % world_size ranks get one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """rows of `t` (labels / seeds) owned by `rank`"""
    a, b = shard_bounds(t.shape[0], world_size, rank)
    return t[a:b]


def gather_to(t: torch.Tensor, n_total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank shards of a tensor (in rank order) on `dst`; returns None on the other ranks.
    Shards may have different lengths (n_total not divisible by the world size)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs: Optional[List[torch.Tensor]] = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)])


def generate_sharded(generator, labels: torch.Tensor, seeds: torch.Tensor, dst: int = 0, **kw):
    """Every rank samples its shard of (labels, seeds) with `generator.generate_latents`; the latents are gathered on
    `dst`.  `labels`/`seeds` are the FULL batch on every rank (they are tiny: 768 + 4*h*w floats per image)."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lab, sd = shard(labels, world, rank), shard(seeds, world, rank)
    lat = generator.generate_latents(lab, num_imgs=lab.shape[0], seeds=sd, **kw) if lab.shape[0] else seeds[:0]
    return gather_to(lat, labels.shape[0], dst=dst)
