"""Host logic of the training exchange step (x3d_tf_b200/exchange.py) under a world-size-2 gloo
group on CPU: bucket layout, asynchronous per-bucket sum all-reduce started in backward order,
finish() semantics, BN moving-statistics averaging.  The NCCL run of the same code is
tests/test_gpu_multi.py (two GPUs); this test needs none."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from x3d_tf_b200.exchange import GradientExchange, average_, make_buckets


def test_buckets_tile_the_arena_in_backward_order():
    b = make_buckets(1000, [300, 700])
    assert b == [(700, 1000), (300, 700), (0, 300)]
    assert make_buckets(10, []) == [(0, 10)]
    assert make_buckets(10, [0, 10, 4, 4]) == [(4, 10), (0, 4)]
    flat = torch.zeros(10)
    with pytest.raises(ValueError):
        GradientExchange(flat, [(0, 4), (5, 10)], 1)           # gap
    with pytest.raises(ValueError):
        GradientExchange(flat, [(0, 4), (4, 9)], 1)            # short
    ex = GradientExchange(flat, make_buckets(10, [4]), 1)
    ex.start(0)
    with pytest.raises(RuntimeError):
        ex.start(0)                                            # a bucket is exchanged once per step
    ex.finish()
    ex.start(0)                                                # next step
    ex.finish()


def test_trainer_bucket_edges_follow_the_stage_layout():
    """The trainer's three buckets: [stage 5 + conv5 + fc1 + fc2], [stage 4], [stem + stages 2, 3] --
    computed from the arena layout only (no GPU needed to check it)."""
    from x3d_tf_b200.arch import build_arch, variable_shapes
    from x3d_tf_b200.config import get_config
    arch = build_arch(get_config("X3D_M"))
    names = [k for k in variable_shapes(arch) if not k.endswith(("moving_mean", "moving_variance"))]
    by_stage = {s: sum(int(np.prod(v)) for k, v in variable_shapes(arch).items()
                       if k.startswith(f"stages/{s}/") and k in names) for s in range(4)}
    total = sum(int(np.prod(variable_shapes(arch)[k])) for k in names)
    tail = total - sum(by_stage[s] for s in range(3)) - sum(
        int(np.prod(variable_shapes(arch)[k])) for k in names if k.startswith("conv1/"))
    assert tail / total > 0.8 and by_stage[2] / total > 0.1      # what makes early exchange worth it


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        rng = np.random.default_rng(100 + rank)
        mine = torch.from_numpy(rng.normal(size=n).astype(np.float32))
        flat = mine.clone()
        ex = GradientExchange(flat, make_buckets(n, [300, 700]), world)
        # backward order: the head's bucket first, while "backward" still writes the others
        ex.start(0)
        flat[0:300] += 1.0                       # a late gradient contribution to a bucket not yet started
        ex.start(1)
        ex.finish()                              # starts bucket 2 itself, waits for all three
        mov = [torch.full((4,), float(rank)), torch.full((2,), 10.0 * (rank + 1))]
        average_(mov, world)
        # second step reuses the object
        flat2_before = flat.clone()
        ex.finish()
        q.put((rank, mine.numpy(), flat2_before.numpy(), flat.numpy(), [m.numpy() for m in mov]))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, mine, first, second, mov = q.get(timeout=240)
        out[rank] = (mine, first, second, mov)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = out[0][0] + out[1][0]
    want[0:300] += 2.0                                         # both ranks' late contribution
    for r in range(2):
        np.testing.assert_allclose(out[r][1], want, rtol=0, atol=1e-6)     # sum of the shards, every bucket
        np.testing.assert_allclose(out[r][2], 2 * want, rtol=0, atol=1e-5)  # a second exchange sums again
        np.testing.assert_array_equal(out[r][3][0], np.full(4, 0.5, np.float32))
        np.testing.assert_array_equal(out[r][3][1], np.full(2, 15.0, np.float32))
    np.testing.assert_array_equal(out[0][1], out[1][1])        # replicas hold identical gradients
