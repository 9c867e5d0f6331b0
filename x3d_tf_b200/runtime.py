"""`utils.get_strategy` / `utils.get_precision` of the reference (utils.py:144-192) for this build.

The reference picks a `tf.distribute` strategy (one process drives every GPU; MirroredStrategy
splits the batch and all-reduces gradients) and a Keras mixed-precision policy.  Here one process
drives ONE GPU (torchrun starts a rank per GPU, `torch.distributed` over NCCL / NVLink is the
plumbing), so the strategy object only records which replica this process is; and the 16-bit
policy is bfloat16 storage with fp32 accumulation (the tensor-core kernels of this build), not
float16: same activation footprint as the reference's `mixed_float16`, fp32 softmax like
`model.py:109-111`, and no loss scaling is needed (bf16 has fp32's exponent range).
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass

import torch


@dataclass
class Strategy:
    """What `strategy.scope()` / `strategy.num_replicas_in_sync` are used for in train.py:127 / eval.py:74."""
    rank: int
    world: int
    device: torch.device

    @property
    def num_replicas_in_sync(self) -> int:
        return self.world

    @contextlib.contextmanager
    def scope(self):
        prev = torch.cuda.current_device()
        torch.cuda.set_device(self.device)
        try:
            yield self
        finally:
            torch.cuda.set_device(prev)

    def shard(self, n_items: int):
        """[lo, hi) of this replica's contiguous block of `n_items` videos (shard.shard_range)."""
        from .shard import shard_range
        return shard_range(n_items, self.world, self.rank)


def get_strategy(num_gpus: int) -> Strategy:
    """utils.py:144-174.  `num_gpus` > 1 needs that many ranks (torchrun --nproc-per-node num_gpus):
    MirroredStrategy's replicas are processes here.  Unlike the reference there is NO CPU strategy:
    without a CUDA device this raises (the path has no CPU implementation)."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the X3D path has no CPU implementation (the reference falls back "
                           "to OneDeviceStrategy('CPU:0') here)")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    want = max(int(num_gpus), 1)
    if want != world:
        raise RuntimeError(f"num_gpus={num_gpus} needs {want} ranks, one per GPU "
                           f"(python -m torch.distributed.run --nproc-per-node {want} ...); WORLD_SIZE={world}")
    if world > torch.cuda.device_count() and world > 1:
        raise RuntimeError(f"{world} ranks but only {torch.cuda.device_count()} visible GPUs")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=dev)
    return Strategy(rank=rank, world=world, device=dev)


def get_precision(mixed_precision: bool) -> str:
    """utils.py:176-192: 'float32', or the 16-bit policy when asked for and a GPU exists.  The policy
    name says what this build computes in: 'mixed_bfloat16' (see the module docstring)."""
    if mixed_precision and torch.cuda.is_available():
        return "mixed_bfloat16"
    return "float32"


def policy_dtype(precision: str) -> str:
    """The `dtype=` argument of `X3D(cfg, dtype=...)` for a policy name ('mixed_float16' is accepted
    for source compatibility and means the same 16-bit policy)."""
    if precision in ("mixed_bfloat16", "mixed_float16"):
        return "bfloat16"
    if precision == "float32":
        return "float32"
    raise ValueError(f"unknown precision policy {precision!r}")
