"""The exchange step of data-parallel training: sum all-reduce of the flat gradient arena.

Reference: `tf.distribute.MirroredStrategy` (`utils.py:160-167`) sums the replicas' gradients with
one NCCL all-reduce inside Keras `fit` (`train.py:145-152`).  Here every rank holds the same flat
fp32 arena (`training._Arena`); the arena is cut into a few contiguous buckets in BACKWARD order --
the head and the last stages hold 97 % of the parameters and their gradients are complete first --
and each bucket's all-reduce is started (asynchronously, NCCL over NVLink) as soon as the backward
pass has produced it, so the exchange overlaps the rest of the backward; `finish()` joins them
before the optimizer step.  The loss is pre-scaled by 1/world, so the sum is the global-batch mean.

No CUDA dependency: the same code runs under `gloo` on CPU tensors (tests/test_exchange.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch


def make_buckets(size: int, edges: Sequence[int]) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) ranges that tile [0, size), cut at `edges` (offsets inside the arena),
    listed in backward order (highest offsets first: the arena is laid out in forward order)."""
    cuts = sorted({0, int(size), *[int(e) for e in edges if 0 < int(e) < size]})
    return [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])][::-1]


class GradientExchange:
    """`start(k)` launches the all-reduce of bucket k (backward order), `finish()` waits for all of
    them.  With world == 1 both are no-ops."""

    def __init__(self, flat: torch.Tensor, buckets: List[Tuple[int, int]], world: int, group=None):
        covered = sorted(buckets)
        if not covered or covered[0][0] != 0 or covered[-1][1] != flat.numel() or \
                any(a[1] != b[0] for a, b in zip(covered[:-1], covered[1:])):
            raise ValueError(f"buckets {buckets} do not tile an arena of {flat.numel()} elements")
        self.flat, self.buckets, self.world, self.group = flat, list(buckets), int(world), group
        self._work: List = []
        self._started: set = set()

    def start(self, k: int) -> None:
        if k in self._started:
            raise RuntimeError(f"bucket {k} exchanged twice in one step")
        self._started.add(k)
        if self.world <= 1:
            return
        import torch.distributed as dist
        lo, hi = self.buckets[k]
        self._work.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                          async_op=True))

    def finish(self) -> None:
        """Starts whatever has not been started (in backward order) and waits for every bucket."""
        for k in range(len(self.buckets)):
            if k not in self._started:
                self.start(k)
        for w in self._work:
            w.wait()
        self._work.clear()
        self._started.clear()


def average_(tensors: Iterable[torch.Tensor], world: int, group=None) -> None:
    """In-place mean over ranks (BN moving statistics when a checkpoint is written: per-replica
    BatchNormalization keeps them local during training)."""
    if world <= 1:
        return
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= world
