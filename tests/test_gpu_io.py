"""GPU parity of the steps either side of the forward path:
uint8 input stage (utils.py:42-72), evaluation metrics (eval.py:62-70) and `X3D.evaluate`."""
import numpy as np
import pytest
import torch

from oracle import io_oracle as IO
from oracle import x3d_oracle as O
from tests.gpu_util import dev, to_dev, to_np
from x3d_tf_b200.arch import build_arch
from x3d_tf_b200.config import get_config
from x3d_tf_b200.synth import synthetic_clips_u8, synthetic_weights

pytestmark = pytest.mark.gpu
MEAN, STD = [0.433, 0.404, 0.377], [0.151, 0.148, 0.157]


def _ops():
    from x3d_tf_b200 import ops
    return ops


@pytest.mark.parametrize("shape", [(2, 3, 8, 8, 3), (1, 1, 5, 7, 3), (1, 2, 3, 3, 3), (3, 4, 33, 17, 3)])
def test_normalize_u8_bit_exact(shape):
    """integer in, IEEE fp32 out in the reference's operation order: bit-exact against numpy."""
    u8 = np.random.default_rng(sum(shape)).integers(0, 256, size=shape, dtype=np.uint8)
    if u8.size >= 512:
        u8.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)         # every pixel value at least once
    want = IO.normalize(u8, MEAN, STD)
    x = torch.from_numpy(u8).to(dev())
    got = _ops().normalize_u8(x, MEAN, STD, torch.float32)
    assert np.array_equal(to_np(got).view(np.uint32), want.view(np.uint32))
    got16 = _ops().normalize_u8(x, MEAN, STD, torch.bfloat16)
    want16 = torch.from_numpy(want).to(torch.bfloat16)
    assert torch.equal(got16.cpu(), want16)


@pytest.mark.parametrize("N,T,H,W", [(2, 4, 32, 32), (1, 5, 23, 37), (1, 1, 16, 16)])
def test_stem_u8_fused_equals_normalize_then_stem(N, T, H, W):
    ops = _ops()
    from tests.test_gpu_ops import _pack_stem_tc
    rng = np.random.default_rng(N + T + H + W)
    u8 = rng.integers(0, 256, size=(N, T, H, W, 3), dtype=np.uint8)
    C = 24
    ks = rng.normal(size=(1, 3, 3, 3, C)).astype(np.float32) * 0.3
    kt = rng.normal(size=(5, 1, 1, 1, C)).astype(np.float32) * 0.5
    bias = to_dev(rng.normal(size=C).astype(np.float32) * 0.1)
    wc = _pack_stem_tc(ks, kt, C)
    x = torch.from_numpy(u8).to(dev())
    fused = ops.stem_tc_u8_fwd(x, MEAN, STD, wc, bias)
    unfused = ops.stem_tc_fwd(ops.normalize_u8(x, MEAN, STD, torch.bfloat16), wc, bias)
    also = ops.stem_tc_fwd(ops.normalize_u8(x, MEAN, STD, torch.float32), wc, bias)
    torch.cuda.synchronize()
    assert torch.equal(fused, unfused) and torch.equal(fused, also)


@pytest.mark.parametrize("V,ncls,k", [(4, 7, 5), (37, 400, 5), (1, 400, 1), (130, 11, 3)])
def test_eval_metrics_match_oracle(V, ncls, k):
    rng = np.random.default_rng(V * 1000 + ncls)
    logits = rng.normal(size=(V, ncls)).astype(np.float32) * 3
    p = np.exp(logits - logits.max(1, keepdims=True)); p = (p / p.sum(1, keepdims=True)).astype(np.float32)
    labels = rng.integers(0, ncls, size=V).astype(np.int32)
    if V >= 4:                                         # ties, exact zeros and ones
        p[0] = 0; p[0, :4] = 0.25; labels[0] = 1
        p[1] = 0; p[1, 2] = 1.0; labels[1] = 0
        labels[2] = int(p[2].argmax())
    want = IO.eval_metrics(p, labels, k)
    acc = torch.zeros(4, dtype=torch.float64, device=dev())
    _ops().eval_metrics(torch.from_numpy(p).to(dev()), torch.from_numpy(labels).to(dev()), acc, k)
    _ops().eval_metrics(torch.from_numpy(p).to(dev()), torch.from_numpy(labels).to(dev()), acc, k)   # accumulates
    got = acc.cpu().numpy() / 2
    assert got[1] == want["sums"][1] and got[2] == want["sums"][2] and got[3] == V       # counts: exact
    assert abs(got[0] - want["sums"][0]) <= 1e-5 * abs(want["sums"][0]) + 1e-6           # fp32 logs


def _model(variant, dtype, views):
    from x3d_tf_b200 import model as M
    M.reset_block_counters()
    cfg = get_config(variant, freeze=False)
    cfg.TEST.NUM_TEMPORAL_VIEWS, cfg.TEST.NUM_SPATIAL_CROPS = views, 1
    cfg.freeze()
    m = M.X3D(cfg, dtype=dtype, use_cuda_graph=True)
    W = synthetic_weights(build_arch(cfg))
    m.set_weights_dict(W)
    return m, cfg, W


@pytest.mark.parametrize("dtype,stem_u8", [("float32", "normalize"), ("bfloat16", "normalize"),
                                           ("bfloat16", "fused")])
def test_model_uint8_clips_equal_normalised_float_clips(dtype, stem_u8, monkeypatch):
    """uint8 clips through the on-device input stage (either form) give bit-identical logits to
    float32 clips normalised on the host the way dataloader.py does."""
    from x3d_tf_b200 import model as M
    monkeypatch.setattr(M.Options, "stem_u8", stem_u8)
    m, cfg, W = _model("X3D_XS", dtype, 2)
    u8 = synthetic_clips_u8(4, 4, 64, 64, seed=9)
    xf = IO.normalize(u8, cfg.DATA.MEAN, cfg.DATA.STD)
    p_u8 = m(torch.from_numpy(u8).to(dev())).clone(); l_u8 = m.last_logits.clone()
    p_f = m(to_dev(xf)).clone(); l_f = m.last_logits.clone()
    torch.cuda.synchronize()
    assert torch.equal(l_u8, l_f) and torch.equal(p_u8, p_f)
    want = O.forward(W, O.OracleSpec.from_cfg(cfg), xf, torch.float64)
    err = np.abs(to_np(l_u8) - want["logits"]).max() / np.abs(want["logits"]).max()
    assert err < (1e-4 if dtype == "float32" else 2e-2)


def test_evaluate_matches_oracle_metrics():
    """X3D.evaluate (device-side metrics, pipelined predict) == oracle metrics of the oracle's
    own probabilities, on uint8 batches with a ragged last batch."""
    m, cfg, W = _model("X3D_XS", "float32", 2)
    m.compile(top_k=5)
    rng = np.random.default_rng(5)
    batches, all_clips, all_labels = [], [], []
    for nv in (3, 3, 2):
        u8 = rng.integers(0, 256, size=(nv * 2, 4, 64, 64, 3), dtype=np.uint8)
        lab = rng.integers(0, 400, size=nv).astype(np.int32)
        batches.append((u8, lab)); all_clips.append(u8); all_labels.append(lab)
    xf = IO.normalize(np.concatenate(all_clips), cfg.DATA.MEAN, cfg.DATA.STD)
    want_p = O.forward(W, O.OracleSpec.from_cfg(cfg), xf, torch.float64)["probs"]
    labels = np.concatenate(all_labels)
    labels[0] = int(want_p[0].argmax())                 # at least one hit
    batches[0] = (batches[0][0], labels[:3])
    want = IO.eval_metrics(want_p.astype(np.float32), labels, 5)
    got = m.evaluate(batches)
    assert got["videos"] == 8
    assert got["acc"] == want["acc"] and got["top_5_acc"] == want["top_5_acc"]
    assert abs(got["loss"] - want["loss"]) < 1e-3 * abs(want["loss"])
    with pytest.raises(ValueError):
        m.evaluate([(batches[0][0], labels[:2])])       # label count must match the videos


def test_eval_driver_finds_checkpoint_and_evaluates(tmp_path, capsys):
    """`python -m x3d_tf_b200.eval` flow (eval.py:26-91): `checkpoint` file -> newest bundle ->
    load_weights(...).expect_partial() -> evaluate.  The bundle is written by X3D.save_weights with
    seeded weights; the driver's metrics must equal a direct evaluate() of the same model on the same
    (seeded, per-video) synthetic set, and differ from a randomly re-initialised model's."""
    from x3d_tf_b200 import eval as E
    from x3d_tf_b200 import model as M
    cfg = get_config("X3D_XS")
    M.reset_block_counters()
    m = M.X3D(cfg, dtype="bfloat16")
    m.set_weights_dict(synthetic_weights(build_arch(cfg), seed=77))
    m.save_weights(str(tmp_path / "ckpt-3"))
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "ckpt-3"\nall_model_checkpoint_paths: "ckpt-3"\n')
    num_preds = cfg.TEST.NUM_TEMPORAL_VIEWS * cfg.TEST.NUM_SPATIAL_CROPS
    vpb = max(cfg.TEST.BATCH_SIZE // num_preds, 1)
    want = m.compile().evaluate(E.synthetic_batches(0, 3, vpb, num_preds, cfg.DATA.TEMP_DURATION,
                                                    cfg.DATA.TEST_CROP_SIZE, cfg.NETWORK.NUM_CLASSES))
    got = E.run(["--cfg", "X3D_XS", "--model_folder", str(tmp_path), "--synthetic", "3"])
    assert got == want and got["videos"] == 3
    assert "top_5_acc" in capsys.readouterr().out
    # a folder without a checkpoint: the reference logs 'No checkpoint found!' and evaluates nothing
    empty = tmp_path / "empty"
    empty.mkdir()
    assert E.run(["--cfg", "X3D_XS", "--model_folder", str(empty), "--synthetic", "3"]) is None


@pytest.mark.parametrize("F,H,W,T,views,crops,S", [(40, 20, 28, 4, 2, 3, 20), (7, 33, 21, 4, 3, 3, 21),
                                                   (100, 16, 16, 13, 1, 1, 16), (5, 24, 19, 16, 10, 1, 18),
                                                   (64, 182, 242, 13, 2, 3, 182)])
def test_eval_views_u8_is_bit_exact(F, H, W, T, views, crops, S):
    """x3d_eval_views_u8 == the reference's temporal views + uniform crops + batch reshape
    (transforms.py:48-65, 149-222; dataloader.py:107-116), byte for byte: short videos that loop,
    more views than frames, landscape and portrait frames, one and three crops, odd row lengths."""
    from oracle import io_oracle
    from x3d_tf_b200 import ops
    rng = np.random.default_rng(F + H + W)
    video = rng.integers(0, 256, size=(F, H, W, 3), dtype=np.uint8)
    want = io_oracle.eval_views(video, T, views, crops, S)
    got = ops.eval_views_u8(torch.from_numpy(video).cuda(), T, views, crops, S)
    torch.cuda.synchronize()
    assert got.shape == want.shape and np.array_equal(got.cpu().numpy(), want)
