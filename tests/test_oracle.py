"""The torch-CPU oracle against the independent numpy statement of the non-default TF layer
semantics, plus internal consistency of the full forward."""
import numpy as np
import pytest
import torch

from oracle import np_ops
from oracle import x3d_oracle as O
from x3d_tf_b200.arch import build_arch
from x3d_tf_b200.config import get_config
from x3d_tf_b200.synth import synthetic_clips, synthetic_weights


def _ncdhw(a):
    return torch.from_numpy(a).permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize("H,W,stride", [(8, 8, 1), (7, 9, 1), (8, 8, 2), (7, 9, 2), (23, 12, 2),
                                        (5, 4, 2), (1, 1, 1), (2, 3, 2)])
def test_channelwise_same_matches_numpy(H, W, stride):
    rng = np.random.default_rng(H * 100 + W * 10 + stride)
    x = rng.normal(size=(2, 3, H, W, 6))
    k = rng.normal(size=(3, 3, 3, 1, 6))
    want = np_ops.channelwise_conv_same(x, k, stride)
    got = O.conv3d_same(_ncdhw(x), O._k(k, torch.float64), (1, stride, stride), 6)
    assert got.shape[2:] == (3, -(-H // stride), -(-W // stride))
    np.testing.assert_allclose(O.to_ndhwc(got), want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("H,W", [(8, 8), (9, 7), (13, 6)])
def test_stem_matches_numpy(H, W):
    rng = np.random.default_rng(H + W)
    x = rng.normal(size=(2, 4, H, W, 3))
    Wt = {"conv1/conv_s/kernel": rng.normal(size=(1, 3, 3, 3, 8)).astype(np.float32),
          "conv1/conv_t/kernel": rng.normal(size=(5, 1, 1, 1, 8)).astype(np.float32),
          "conv1/bn/gamma": np.ones(8, np.float32), "conv1/bn/beta": np.zeros(8, np.float32),
          "conv1/bn/moving_mean": np.zeros(8, np.float32),
          "conv1/bn/moving_variance": np.ones(8, np.float32) - 1e-5}

    class S:
        temp_filter, conv1_dim, bn_eps = 5, 8, 1e-5
    got = O.to_ndhwc(O.stem(Wt, _ncdhw(x), S, torch.float64))
    want = np.maximum(np_ops.stem_convs(x, Wt["conv1/conv_s/kernel"], Wt["conv1/conv_t/kernel"]), 0)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)   # BN var=1-eps -> scale 1 (f32 rounding)


def test_shortcut_valid_stride_matches_numpy():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(1, 2, 7, 9, 5))
    k = rng.normal(size=(1, 1, 1, 5, 4))
    got = O.to_ndhwc(O.conv3d_valid(_ncdhw(x), O._k(k, torch.float64), (1, 2, 2)))
    np.testing.assert_allclose(got, np_ops.pointwise_conv_valid(x, k, 2), rtol=1e-12, atol=1e-12)
    assert got.shape == (1, 2, 4, 5, 4)


def test_forward_view_average_and_softmax():
    cfg = get_config("X3D_XS", freeze=False)
    cfg.TEST.NUM_TEMPORAL_VIEWS = 2
    arch = build_arch(cfg)
    W = synthetic_weights(arch, seed=5)
    spec = O.OracleSpec.from_cfg(cfg)
    x = synthetic_clips(4, 4, 32, 32, cfg.DATA.MEAN, cfg.DATA.STD, seed=2)
    r = O.forward(W, spec, x, torch.float64)
    assert r["logits"].shape == (4, 400) and r["probs"].shape == (2, 400)
    p = np.exp(r["logits"] - r["logits"].max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    np.testing.assert_allclose(r["clip_probs"], p, rtol=1e-10)
    np.testing.assert_allclose(r["probs"], p.reshape(2, 2, 400).mean(1), rtol=1e-10)
    np.testing.assert_allclose(r["probs"].sum(-1), 1.0, rtol=1e-10)
    # each clip is independent of its batch neighbours (inference BN, per-clip SE)
    r1 = O.forward(W, spec, x[2:4], torch.float64)
    np.testing.assert_allclose(r1["logits"], r["logits"][2:4], rtol=1e-9, atol=1e-12)
    # fp32 evaluation agrees with fp64 far inside the 1e-4 contract
    r32 = O.forward(W, spec, x, torch.float32)
    assert np.abs(r32["logits"] - r["logits"]).max() / np.abs(r["logits"]).max() < 1e-5
    with pytest.raises(AssertionError):
        O.forward(W, spec, x[:3], torch.float64)      # batch not a multiple of num_preds


def test_se_taps_only_on_se_blocks():
    cfg = get_config("X3D_XS")
    W = synthetic_weights(build_arch(cfg), seed=1)
    spec = O.OracleSpec.from_cfg(cfg)
    taps = {}
    O.forward(W, spec, synthetic_clips(1, 4, 32, 32, cfg.DATA.MEAN, cfg.DATA.STD), torch.float32,
              training=True, taps=taps)
    se = sorted(k for k in taps if k.endswith("/se_scale"))
    assert len(se) == 13
    assert se[0] == "stages/0/stage/layer_with_weights-0/bottleneck/se_scale"


def test_tf32x3_split_keeps_fp32_accuracy():
    """The 3xTF32 split behind x3d_pw_fwd / x3d_pw_wgrad (fp32): the arithmetic model in np_ops bounds
    its error at ~2^-21 of |a|.|b| per product (lo.lo dropped, lo rounded to TF32), i.e. below the
    rounding of an fp32 GEMM of the layer shapes; a plain TF32 product (hi.hi only) is ~1000x worse,
    which is why the single-MMA form cannot hold the 1e-4 logit bound of the fp32 path."""
    from oracle import np_ops
    rng = np.random.default_rng(0)
    for K, N in ((24, 56), (216, 96), (432, 192)):
        a = rng.normal(size=(256, K)).astype(np.float32)
        b = (rng.normal(size=(K, N)) * 0.1).astype(np.float32)
        want = a.astype(np.float64) @ b.astype(np.float64)
        scale = np.abs(want).max()
        err3 = np.abs(np_ops.tf32x3_matmul(a, b) - want).max() / scale
        fp32 = np.abs((a @ b).astype(np.float64) - want).max() / scale
        hi = lambda x: ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
        err1 = np.abs(hi(a) @ hi(b) - want).max() / scale
        assert err3 < 2e-6 and err3 < 4 * max(fp32, 1e-7), (K, N, err3, fp32)
        assert err1 > 50 * err3, (K, N, err1, err3)
    # exactness of the split itself: hi + lo == x, hi has 10 mantissa bits
    x = rng.normal(size=4096).astype(np.float32)
    bits = x.view(np.uint32)
    h = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    assert np.array_equal((h.astype(np.float64) + (x - h).astype(np.float64)).astype(np.float32), x)
    assert not (h.view(np.uint32) & np.uint32(0x1FFF)).any()
