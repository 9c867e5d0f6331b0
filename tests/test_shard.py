"""Multi-GPU inference shards whole videos across ranks with no data-path collective
(SURVEY.md section 8e).  The partitioning and the result collection are exercised with a
world-size-2 gloo group on CPU; the per-rank compute is stood in by the CPU oracle (test
infrastructure), which is batch-independent like the CUDA path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from x3d_tf_b200.shard import gather_predictions, shard_clips, shard_range


@pytest.mark.parametrize("n,world", [(8, 1), (8, 2), (7, 2), (3, 4), (0, 2), (17, 8)])
def test_shard_range_tiles_the_videos(n, world):
    blocks = [shard_range(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a, b), (c, d) in zip(blocks, blocks[1:]):
        assert b == c
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1


def test_shard_clips_keeps_views_together():
    clips = np.arange(6 * 2).reshape(12, 1, 1, 1, 1)          # 6 videos x 2 views
    a, b = shard_clips(clips, 2, 2, 0), shard_clips(clips, 2, 2, 1)
    assert a.shape[0] == 6 and b.shape[0] == 6 and a[0, 0, 0, 0, 0] == 0 and b[0, 0, 0, 0, 0] == 6
    with pytest.raises(ValueError):
        shard_clips(clips[:5], 2, 2, 0)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import x3d_oracle as O
        from x3d_tf_b200.arch import build_arch
        from x3d_tf_b200.config import get_config
        from x3d_tf_b200.synth import synthetic_clips, synthetic_weights
        torch.set_num_threads(2)
        cfg = get_config("X3D_XS", freeze=False)
        cfg.TEST.NUM_TEMPORAL_VIEWS = 2
        cfg.freeze()
        W = synthetic_weights(build_arch(cfg))
        videos, views = 3, 2                                   # uneven split: 2 + 1 videos
        clips = synthetic_clips(videos * views, 4, 32, 32, cfg.DATA.MEAN, cfg.DATA.STD, seed=3)
        spec = O.OracleSpec.from_cfg(cfg)
        mine = shard_clips(clips, views, world, rank)
        local = torch.from_numpy(O.forward(W, spec, mine, torch.float64)["probs"])
        full = gather_predictions(local, videos)
        if rank == 0:
            want = O.forward(W, spec, clips, torch.float64)["probs"]
            q.put(float(np.abs(full.numpy() - want).max()))
    finally:
        dist.destroy_process_group()


def test_sharded_eval_equals_single_process_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12
