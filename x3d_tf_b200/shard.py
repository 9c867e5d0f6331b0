"""Video sharding for multi-GPU inference (one process per GPU, no data-path collective).

The reference splits each eval batch across GPUs with `tf.distribute.MirroredStrategy`
(`utils.py:160-167`, `eval.py:72-89`).  The unit that must stay on one GPU is a *video*: its
`NUM_TEMPORAL_VIEWS*NUM_SPATIAL_CROPS` consecutive clips are averaged by `X3D.call`
(`model.py:123-126`, layout from `dataloader.py:107-116`).  Videos are independent, so ranks
never exchange activations; only the per-video probabilities are collected at the end.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def shard_range(num_videos: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of videos for `rank`: sizes differ by at most one and the
    blocks tile [0, num_videos) in rank order."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(num_videos, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_clips(clips, num_preds: int, world: int, rank: int):
    """Slice of an eval batch `[videos*num_preds, T, H, W, C]` owned by `rank` (whole videos)."""
    n = clips.shape[0]
    if n % num_preds:
        raise ValueError(f"batch {n} is not a multiple of num_preds {num_preds}")
    lo, hi = shard_range(n // num_preds, world, rank)
    return clips[lo * num_preds:hi * num_preds]


def gather_predictions(local: torch.Tensor, num_videos: int, group=None) -> torch.Tensor:
    """Collects per-video rows `[local_videos, classes]` from every rank into
    `[num_videos, classes]` in video order (ranks own contiguous blocks, see `shard_range`).
    Works for CPU (gloo) and CUDA (nccl) tensors; with no process group returns `local`."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(num_videos, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((most, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
