"""Two-GPU tests (NCCL over NVLink): the sharded evaluation and the data-parallel training step.
Skipped on a box with fewer than two GPUs (the CPU suite covers the same host logic with gloo:
tests/test_shard.py, tests/test_io_oracle.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _spawn(fn, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = {}
    import queue as _queue
    import time as _time
    deadline = _time.time() + 300
    while len(out) < world:
        try:
            k, v = q.get(timeout=5)
            out[k] = v
        except _queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead or _time.time() > deadline:          # a crashed rank must fail the test, not hang it
                [p.kill() for p in procs if p.is_alive()]
                raise AssertionError(f"worker exit codes {[p.exitcode for p in procs]}")
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return out


def _init(rank, world, port):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    return dist


def _eval_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    from x3d_tf_b200 import eval as E
    from x3d_tf_b200 import model as M
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.shard import shard_range
    from x3d_tf_b200.synth import synthetic_weights
    cfg = get_config("X3D_XS", freeze=False)
    cfg.TEST.NUM_TEMPORAL_VIEWS, cfg.TEST.NUM_SPATIAL_CROPS = 2, 1
    cfg.freeze()
    M.reset_block_counters()
    m = M.X3D(cfg, dtype="bfloat16").compile()
    m.set_weights_dict(synthetic_weights(build_arch(cfg)))
    V = 7
    lo, hi = shard_range(V, world, rank)
    res = m.evaluate(E.synthetic_batches(lo, hi, 2, 2, 4, 64, 400))          # all-reduced over ranks
    dist.barrier()
    dist.destroy_process_group()
    alone = m.evaluate(E.synthetic_batches(0, V, 2, 2, 4, 64, 400)) if rank == 0 else None
    q.put((rank, (res, alone)))


def test_sharded_evaluate_equals_single_gpu():
    _need_two()
    out = _spawn(_eval_worker)
    (r0, alone), (r1, _) = out[0], out[1]
    assert r0 == r1 and r0["videos"] == 7
    assert r0["acc"] == alone["acc"] and r0["top_5_acc"] == alone["top_5_acc"]
    assert abs(r0["loss"] - alone["loss"]) < 1e-9 * max(1.0, abs(alone["loss"]))


def _train_worker(rank, world, port, q):
    dist = _init(rank, world, port)
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.synth import synthetic_clips, synthetic_weights
    from x3d_tf_b200.training import X3DTrainer
    cfg = get_config("X3D_XS", freeze=False)
    cfg.NETWORK.DROPOUT_RATE = 0.0
    cfg.freeze()
    W = synthetic_weights(build_arch(cfg), seed=3)
    xs = [synthetic_clips(2, 4, 64, 64, cfg.DATA.MEAN, cfg.DATA.STD, seed=10 + r) for r in range(world)]
    ls = [np.random.default_rng(20 + r).integers(0, 400, size=2).astype(np.int32) for r in range(world)]
    dev = torch.device("cuda", rank)
    tr = X3DTrainer(cfg, device=dev, world=world).load(W)
    tr.step(torch.from_numpy(xs[rank]).to(dev), torch.from_numpy(ls[rank]).to(dev), 0.05)
    torch.cuda.synchronize()
    got_g, got_w = tr.g.clone(), tr.w.clone()
    dist.barrier()
    dist.destroy_process_group()
    ref = None
    if rank == 0:
        # the same exchange done by hand on one GPU: each shard's gradients (already scaled by
        # 1/world) summed, then the identical update
        parts = []
        for r in range(world):
            t1 = X3DTrainer(cfg, device=dev, world=world).load(W)
            t1.forward_backward(torch.from_numpy(xs[r]).to(dev), torch.from_numpy(ls[r]).to(dev))
            parts.append(t1.g64.clone())
        ref = (parts[0] + parts[1]).to(torch.float32).cpu().numpy()
    q.put((rank, (got_g.cpu().numpy(), got_w.cpu().numpy(), ref)))


def test_training_step_allreduce_two_gpus():
    _need_two()
    out = _spawn(_train_worker)
    g0, w0, ref = out[0]
    g1, w1, _ = out[1]
    assert np.array_equal(g0, g1) and np.array_equal(w0, w1)          # replicas stay identical
    scale = np.abs(ref).max()
    assert np.abs(g0 - ref).max() <= 2e-6 * scale + 1e-12
