"""CPU tests of the input-stage / eval-metric oracle, the eval driver's host logic and the
world-size-2 metric reduction (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import io_oracle as IO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_normalize_known_answers():
    mean, std = [0.433, 0.404, 0.377], [0.151, 0.148, 0.157]          # configs/kinetics/X3D_M.yaml:17-18
    u8 = np.array([[[[[0, 128, 255]]]]], np.uint8)
    got = IO.normalize(u8, mean, std)
    want = [(0 / 255 - 0.433) / 0.151, (128 / 255 - 0.404) / 0.148, (1.0 - 0.377) / 0.157]
    np.testing.assert_allclose(got.reshape(-1), want, rtol=3e-7)
    assert got.dtype == np.float32
    # range of normalised pixels: [(0 - 0.433)/0.151, (1 - 0.404)/0.148]
    full = IO.normalize(np.array([[0, 0, 0], [255, 255, 255]], np.uint8), mean, std)
    assert abs(full.min() + 2.8675) < 1e-3 and abs(full.max() - 4.0270) < 1e-3
    # the product's synthetic-clip generator follows the same definition
    from x3d_tf_b200.synth import normalize_clips
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(2, 3, 5, 7, 3), dtype=np.uint8)
    assert np.array_equal(normalize_clips(x, mean, std), IO.normalize(x, mean, std))


def test_eval_metrics_hand_worked():
    p = np.array([[0.7, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0],       # label 0: top-1 hit
                  [0.1, 0.1, 0.3, 0.2, 0.15, 0.1, 0.05],     # label 6: 6 classes larger -> top-5 miss
                  [0.25, 0.25, 0.25, 0.25, 0.0, 0.0, 0.0],   # label 1: tie, argmax picks index 0 -> top-1 miss, top-5 hit
                  [0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0]],      # label 0: probability 0 -> clipped to 1e-7
                 np.float32)
    lab = np.array([0, 6, 1, 0])
    r = IO.eval_metrics(p, lab, k=5)
    assert r["acc"] == 0.25 and r["top_5_acc"] == 0.75 and r["videos"] == 4   # video 3: one class is larger
    want = [-np.log(0.7), -np.log(0.05), -np.log(0.25), -np.log(1e-7)]
    got_sum = r["sums"][0]
    assert abs(got_sum - sum(want)) < 1e-4 * sum(want)
    # in_top_k: exactly k-1 strictly larger classes is still a hit
    p2 = np.array([[0.3, 0.25, 0.2, 0.15, 0.06, 0.04]], np.float32)
    assert IO.eval_metrics(p2, np.array([4]), k=5)["top_5_acc"] == 1.0
    assert IO.eval_metrics(p2, np.array([5]), k=5)["top_5_acc"] == 0.0


def test_file_list_and_batches(tmp_path):
    from x3d_tf_b200 import eval as E
    for i in range(3):
        np.save(tmp_path / f"v{i}.npy", np.full((2, 1, 4, 4, 3), i, np.uint8))
    lst = tmp_path / "test.txt"
    lst.write_text("# comment\nv0.npy 7\n\nv1.npy 3\n" + str(tmp_path / "v2.npy") + " 11\n")
    items = E.read_file_list(str(lst))
    assert [lab for _, lab in items] == [7, 3, 11] and all(os.path.isabs(p) for p, _ in items)
    batches = list(E.file_batches(items, 2, 2))
    assert [b[0].shape[0] for b in batches] == [4, 2]
    assert batches[0][1].tolist() == [7, 3] and batches[1][0][0, 0, 0, 0, 0] == 2
    with pytest.raises(ValueError):
        list(E.file_batches(items, 2, 3))                     # wrong number of views per video
    bad = tmp_path / "bad.txt"
    bad.write_text("only_a_path\n")
    with pytest.raises(ValueError):
        E.read_file_list(str(bad))


def test_synthetic_batches_do_not_depend_on_sharding():
    from x3d_tf_b200 import eval as E
    from x3d_tf_b200.shard import shard_range
    whole = list(E.synthetic_batches(0, 5, 2, 2, 1, 4, 400))
    clips = np.concatenate([c for c, _ in whole]); labels = np.concatenate([l for _, l in whole])
    parts_c, parts_l = [], []
    for r in range(2):
        lo, hi = shard_range(5, 2, r)
        for c, l in E.synthetic_batches(lo, hi, 3, 2, 1, 4, 400):
            parts_c.append(c); parts_l.append(l)
    assert np.array_equal(np.concatenate(parts_c), clips) and np.array_equal(np.concatenate(parts_l), labels)


def test_eval_driver_requires_gpu_and_checkpoint_dir(tmp_path):
    from x3d_tf_b200 import eval as E
    with pytest.raises(NotADirectoryError):
        E.run(["--cfg", "X3D_XS", "--model_folder", str(tmp_path / "missing"), "--synthetic", "2"])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU implementation"):
            E.run(["--cfg", "X3D_XS", "--model_folder", str(tmp_path), "--synthetic", "2"])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _metric_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from x3d_tf_b200.model import finalize_metrics
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    acc = torch.tensor([[2.0, 1.0, 2.0, 3.0], [4.5, 2.0, 2.0, 2.0]][rank], dtype=torch.float64)
    q.put((rank, finalize_metrics(acc, None, 5)))
    dist.destroy_process_group()


def test_metric_sums_reduce_over_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_metric_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for r in range(2):
        assert res[r]["videos"] == 5
        assert abs(res[r]["loss"] - 6.5 / 5) < 1e-12 and res[r]["acc"] == 0.6 and res[r]["top_5_acc"] == 0.8


def test_eval_views_oracle_hand_worked():
    """transforms.py:48-65 / 149-222 on a video whose pixel value encodes (frame, y, x)."""
    from oracle import io_oracle
    F, H, W = 5, 4, 6
    video = np.zeros((F, H, W, 1), np.int64)
    for f in range(F):
        for y in range(H):
            for x in range(W):
                video[f, y, x, 0] = f * 100 + y * 10 + x
    # T=2, 3 views: rate = max(1, 5 // 2) = 2; frames (k*2) mod 5 for k = 0..5 -> 0,2,4,1,3,0
    out = io_oracle.eval_views(video, T=2, views=3, crops=3, size=4)
    assert out.shape == (9, 2, 4, 4, 1)
    frames = out[:, :, 0, 0, 0] // 100
    assert frames[:3].tolist() == [[0, 2], [4, 1], [3, 0]]            # crop 0: views 0..2
    # landscape (W > H): x offsets 0, ceil((6-4)/2) = 1, 6-4 = 2 for left / centre / right; y offset 0
    assert [int(out[c * 3, 0, 0, 0, 0] % 10) for c in range(3)] == [0, 1, 2]
    assert int(out[0, 0, 0, 0, 0] // 10 % 10) == 0
    one = io_oracle.eval_views(video, T=2, views=1, crops=1, size=4)   # single crop = centre
    assert int(one[0, 0, 0, 0, 0] % 10) == 1
    tall = io_oracle.eval_views(video.transpose(0, 2, 1, 3), T=2, views=1, crops=3, size=4)   # H > W: y offsets
    assert [int(tall[c, 0, 0, 0, 0] // 100 * 0 + tall[c, 0, 0, 0, 0] % 10) for c in range(3)] == [0, 1, 2]
