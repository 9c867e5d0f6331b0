"""Block-level and whole-model parity of the CUDA path against the CPU oracle, plus
size-independent properties at larger shapes (tiers 3 and 5)."""
import numpy as np
import pytest
import torch

from oracle import x3d_oracle as O
from tests.gpu_util import assert_close, bf16_round, dev, rel_err, to_dev, to_np
from x3d_tf_b200.arch import build_arch
from x3d_tf_b200.config import get_config
from x3d_tf_b200.synth import synthetic_clips, synthetic_weights

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4       # north_star: fp32 logits within 1e-4 relative
BF16_TOL = 2e-2       # north_star: bf16 logits within 2e-2 relative, identical top-1


def _model(variant, dtype=None, views=None, seed=1111, graph=False):
    from x3d_tf_b200 import model as M
    M.reset_block_counters()
    cfg = get_config(variant, freeze=False)
    if views is not None:
        cfg.TEST.NUM_TEMPORAL_VIEWS, cfg.TEST.NUM_SPATIAL_CROPS = views, 1
    cfg.freeze()
    m = M.X3D(cfg, dtype=dtype, use_cuda_graph=graph)
    # heavy-tailed class scales: the top-1 of (almost) every clip is decided at the error level
    W = synthetic_weights(build_arch(cfg), seed=seed, head_spread=1.5)
    m.set_weights_dict(W)
    return m, cfg, W, O.OracleSpec.from_cfg(cfg)


def _check_logits(got, want, tol, check_top1):
    err = rel_err(got, want)
    assert err < tol, f"relative logit error {err:.3e} >= {tol}"
    if check_top1:
        # identical top-1 wherever the oracle's own margin is resolvable at this error level
        srt = np.sort(want, -1)
        margin = srt[:, -1] - srt[:, -2]
        decided = margin > 2 * np.abs(got - want).max()
        assert decided.mean() >= 0.5, "fixture too flat: top-1 undecided on most clips"
        assert (got.argmax(-1)[decided] == want.argmax(-1)[decided]).all()
    return err


@pytest.mark.parametrize("variant,shape", [("X3D_XS", (4, 4, 64, 64)), ("X3D_XS", (2, 4, 91, 75)),
                                           ("X3D_M", (2, 8, 64, 64))])
def test_whole_model_fp32(variant, shape):
    m, cfg, W, spec = _model(variant, views=2)
    x = synthetic_clips(*shape, cfg.DATA.MEAN, cfg.DATA.STD, seed=3)
    want = O.forward(W, spec, x, torch.float64)
    probs = m(to_dev(x))
    _check_logits(to_np(m.last_logits), want["logits"], FP32_TOL, True)
    np.testing.assert_allclose(to_np(probs), want["probs"], rtol=1e-3, atol=1e-7)
    assert probs.shape == (shape[0] // 2, 400)


@pytest.mark.parametrize("pointwise", ["tc", "simt"])
@pytest.mark.parametrize("variant,shape", [("X3D_XS", (4, 4, 64, 64)), ("X3D_S", (2, 13, 91, 91)),
                                           ("X3D_M", (2, 16, 64, 64))])
def test_whole_model_bf16(variant, shape, pointwise):
    from x3d_tf_b200 import model as M
    M.Options.pointwise = pointwise
    try:
        m, cfg, W, spec = _model(variant, dtype="bfloat16", views=1)
        x = synthetic_clips(*shape, cfg.DATA.MEAN, cfg.DATA.STD, seed=4)
        want = O.forward(W, spec, x, torch.float64)
        m(to_dev(x))                                   # fp32 clips read directly by the stem
        e1 = _check_logits(to_np(m.last_logits), want["logits"], BF16_TOL, True)
        m(to_dev(x, torch.bfloat16))                   # bf16 clips (BASELINE config 2)
        want_b = O.forward(W, spec, bf16_round(x), torch.float64)
        _check_logits(to_np(m.last_logits), want_b["logits"], BF16_TOL, True)
        print(f"{variant} {shape} {pointwise}: bf16 rel logit err {e1:.2e}")
    finally:
        M.Options.pointwise = "tc"


def test_large_variants_build_and_run_bf16():
    """X3D-L / XL graphs (55 blocks, SE on flipped parity, 630-wide stage) at a tiny input."""
    for variant in ("X3D_L", "X3D_XL"):
        m, cfg, W, spec = _model(variant, dtype="bfloat16", views=1)
        x = synthetic_clips(1, 4, 45, 39, cfg.DATA.MEAN, cfg.DATA.STD, seed=5)
        want = O.forward(W, spec, x, torch.float64)
        m(to_dev(x))
        _check_logits(to_np(m.last_logits), want["logits"], BF16_TOL, False)


def test_blocks_standalone_api():
    """Reference class API: ResStage / ResBlock / Bottleneck / X3D_Stem / AdaptiveAvgPool3D."""
    from x3d_tf_b200 import model as M
    M.reset_block_counters()
    cfg = get_config("X3D_M")
    rng = np.random.default_rng(0)
    # ResStage(24 -> 54 -> 24, depth 3): blocks 1..3, SE on 1 and 3; stride-2 shortcut on the first
    st = M.ResStage(in_channels=24, inner_channels=54, out_channels=24, depth=3, bn_cfg=cfg.NETWORK.BN)
    assert st._inner_channels == 54 and [b.bottleneck.has_se for b in st.blocks] == [True, False, True]
    arch = build_arch(cfg)
    W = synthetic_weights(arch, seed=9)
    st.set_weights_dict(W, prefix="stages/0/")
    x = rng.normal(size=(2, 4, 15, 18, 24)).astype(np.float32)
    spec = O.OracleSpec.from_cfg(cfg)
    ref = torch.from_numpy(x).double().permute(0, 4, 1, 2, 3)
    for blk in spec.blocks[:3]:
        ref = O.res_block(W, ref, blk, spec, torch.float64)
    got = st(to_dev(x))
    assert got.shape == (2, 4, 8, 9, 24)
    assert_close(to_np(got), O.to_ndhwc(ref), torch.float32, "ResStage")
    # Bottleneck alone (no residual, no final ReLU)
    bt = st.blocks[1].bottleneck
    y = rng.normal(size=(1, 3, 6, 5, 24)).astype(np.float32)
    refb = O.bottleneck(W, torch.from_numpy(y).double().permute(0, 4, 1, 2, 3),
                        "stages/0/stage/layer_with_weights-1", 54, 1, 0, spec, torch.float64)
    assert_close(to_np(bt(to_dev(y))), O.to_ndhwc(refb), torch.float32, "Bottleneck")
    # pool
    p = M.AdaptiveAvgPool3D((1, 1, 1))(to_dev(y))
    assert p.shape == (1, 1, 1, 1, 24)
    np.testing.assert_allclose(to_np(p).ravel(), y.mean((0, 1, 2, 3)), rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        st(to_dev(x), training=True)


def test_determinism_batch_independence_and_graph_replay():
    """Properties that do not need the oracle: bit-exact reruns, a clip's logits do not depend on
    its batch neighbours (inference BN, per-clip SE), CUDA-graph replay == eager launches."""
    m, cfg, W, spec = _model("X3D_M", dtype="bfloat16", views=1)
    x = to_dev(synthetic_clips(6, 16, 112, 112, cfg.DATA.MEAN, cfg.DATA.STD, seed=8), torch.bfloat16)
    m(x)
    l1 = m.last_logits.clone()
    m(x)
    assert torch.equal(l1, m.last_logits)
    m(x[2:4].contiguous())
    assert torch.equal(l1[2:4], m.last_logits)
    mg, *_ = _model("X3D_M", dtype="bfloat16", views=1, graph=True)
    for _ in range(3):
        mg(x)
    assert torch.equal(l1, mg.last_logits)
    # view averaging: 3 views of the same clip average to that clip's softmax
    m3, *_ = _model("X3D_M", dtype="bfloat16", views=3)
    p3 = m3(x)
    p1 = torch.softmax(l1.float(), -1).reshape(2, 3, -1).mean(1)
    np.testing.assert_allclose(to_np(p3), to_np(p1), rtol=1e-4, atol=1e-8)
    with pytest.raises(ValueError):
        m3(x[:4])


def test_checkpoint_round_trip(tmp_path):
    """save -> TF-bundle on disk -> load_weights(...).expect_partial() -> identical logits."""
    m, cfg, W, _ = _model("X3D_XS", views=1)
    x = to_dev(synthetic_clips(1, 4, 48, 48, cfg.DATA.MEAN, cfg.DATA.STD, seed=2))
    m(x)
    l1 = m.last_logits.clone()
    prefix = str(tmp_path / "ckpt" / "model")
    W2 = dict(W)
    W2["optimizer/iter"] = np.array(7, np.int64)
    from x3d_tf_b200 import tf_bundle
    tf_bundle.write_bundle(prefix, W2)
    m2, *_ = _model("X3D_XS", views=1, seed=999)
    m2(x)
    assert not torch.equal(l1, m2.last_logits)
    status = m2.load_weights(tf_bundle.latest_checkpoint(str(tmp_path / "ckpt")))
    status.expect_partial().assert_existing_objects_matched()
    m2(x)
    assert torch.equal(l1, m2.last_logits)


def test_predict_pipelined_host_batches_equal_call():
    """X3D.predict (pinned host batches, H2D on a copy stream overlapped with the previous
    forward, two graph slots sharing one memory pool) returns exactly what X3D.call returns."""
    m, cfg, W, spec = _model("X3D_XS", dtype="bfloat16", views=2, graph=True)
    batches = [torch.from_numpy(synthetic_clips(4, 4, 64, 64, cfg.DATA.MEAN, cfg.DATA.STD, seed=s))
               .to(torch.bfloat16).pin_memory() for s in range(5)]
    want = [m(b).float().cpu().clone() for b in batches]
    got = list(m.predict(iter(batches)))
    assert len(got) == 5
    for g, w in zip(got, want):
        assert g.shape == (2, 400) and torch.equal(g, w)
    # a second pass reuses the captured slots; numpy float32 input and a different shape also work
    got2 = list(m.predict(b.float().numpy() for b in batches[:3]))
    for g, w in zip(got2, want):
        assert torch.allclose(g, w, atol=2e-3)
    odd = synthetic_clips(2, 4, 46, 38, cfg.DATA.MEAN, cfg.DATA.STD, seed=9)
    (p,) = list(m.predict([odd]))
    ref = O.forward(W, spec, odd, torch.float64)["probs"]
    assert np.abs(p.numpy() - ref).max() < 2e-3
    assert list(m.predict([])) == []
    with pytest.raises(ValueError):
        list(m.predict([odd[:1]]))


@pytest.mark.parametrize("variant,T,S,views,clips", [("X3D_M", 16, 256, 10, 20),      # BASELINE configs[2]
                                                     ("X3D_S", 13, 182, 3, 6),        # configs[1] shape, 3-crop
                                                     ("X3D_L", 16, 356, 1, 3)])       # configs[3] shape
def test_full_size_properties(variant, T, S, views, clips):
    """At BASELINE.json's full clip sizes (where the fp64 oracle is only affordable for a couple of
    clips): two clips against the oracle, then size-independent properties for the whole batch -- a clip's
    logits do not depend on the batch it is in, the video probabilities are the mean of its views'
    softmaxes, rows sum to one, and the graph replay reproduces the eager launches bit for bit."""
    m, cfg, W, spec = _model(variant, dtype="bfloat16", views=views)
    x = synthetic_clips(clips, T, S, S, cfg.DATA.MEAN, cfg.DATA.STD, seed=21)
    xd = to_dev(x, torch.bfloat16)
    probs = m(xd).clone()
    logits = m.last_logits.clone()
    assert torch.isfinite(logits).all() and probs.shape == (clips // views, cfg.NETWORK.NUM_CLASSES)
    # (1) full-size clips against the float64 oracle: the first and the last of the batch (two
    # different videos; one for X3D-L, whose fp64 forward is the slowest), with identical top-1
    spec1 = O.OracleSpec.from_cfg(cfg)
    spec1.num_preds = 1
    pick = [0] if variant == "X3D_L" else [0, clips - 1]
    want = O.forward(W, spec1, x[pick], torch.float64)["logits"]
    _check_logits(to_np(logits[pick]), want, BF16_TOL, True)
    # (2) batch independence: the last `views` clips alone
    m(xd[-views:].contiguous())
    assert torch.equal(m.last_logits, logits[-views:])
    # (3) view mean and normalisation
    ref = torch.softmax(logits.float(), -1).reshape(clips // views, views, -1).mean(1)
    np.testing.assert_allclose(to_np(probs), to_np(ref), rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(to_np(probs).sum(-1), 1.0, rtol=1e-4)
    # (4) CUDA-graph replay == eager
    mg, *_ = _model(variant, dtype="bfloat16", views=views, graph=True)
    mg(xd); mg(xd)
    assert torch.equal(mg.last_logits, logits)


@pytest.mark.parametrize("K,N,residual,swish", [(24, 54, False, False), (54, 24, True, False),
                                                (24, 24, False, False), (54, 24, True, True)])
def test_pixel_pairing_bf16_is_bit_identical(K, N, residual, swish):
    """PointwiseConv's pixel pairing ([M/2, 2K] x blockdiag(W, W)) only adds exact zeros to each
    accumulator, but groups a pixel's products into K=16 MMA steps differently: with and without it
    (factors 2 and 4) the tcgen05 path must agree to fp32-accumulation rounding, i.e. at most one bf16
    ulp on a handful of outputs, including an odd tile count and the residual / swish variants."""
    from x3d_tf_b200 import model as M
    rng = np.random.default_rng(7)
    Mrows = 4 * 1237                                   # divisible by 4, not by the 128-row tile
    ks, ns = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    kern = rng.normal(size=(1, 1, 1, K, N)).astype(np.float32) * 0.2
    scale, shift = rng.uniform(0.5, 1.5, N), rng.normal(size=N) * 0.1
    a = torch.zeros(Mrows, ks, dtype=torch.bfloat16, device=dev())
    a[:, :K] = torch.from_numpy(rng.normal(size=(Mrows, K)).astype(np.float32)).to(dev()).to(torch.bfloat16)
    r = None
    if residual:
        r = torch.zeros(Mrows, ns, dtype=torch.bfloat16, device=dev())
        r[:, :N] = torch.from_numpy(rng.normal(size=(Mrows, N)).astype(np.float32)).to(dev()).to(torch.bfloat16)
    pc = M.PointwiseConv(kern, scale, shift, dev())
    saved = (M.Options.pair_pixels, M.Options.pair_aligned)
    outs = {}
    try:
        for P in (1, 2, 4):
            M.Options.pair_pixels, M.Options.pair_aligned = P, False
            assert pc._pair_factor(Mrows) == P
            outs[P] = pc.run(a, Mrows, use_tc=True, residual=r, swish=swish, relu=True).clone()
        torch.cuda.synchronize()
    finally:
        M.Options.pair_pixels, M.Options.pair_aligned = saved
    assert outs[1].shape == (Mrows, ns)
    for P in (2, 4):
        d = (outs[P].float() - outs[1].float()).abs()
        assert float((d > 0).float().mean()) < 1e-3, (P, float((d > 0).float().mean()))
        assert bool((d <= outs[1].float().abs() * 2.0 ** -7 + 1e-30).all()), P
    want = a[:, :K].double() @ torch.from_numpy(kern.reshape(K, N) * scale[None, :]).to(dev()) + \
        torch.from_numpy(shift).to(dev())
    if swish:
        x = a[:, :K].double()
        want = (x * torch.sigmoid(x)) @ torch.from_numpy(kern.reshape(K, N) * scale[None, :]).to(dev()) + \
            torch.from_numpy(shift).to(dev())
    if residual:
        want = want + r[:, :N].double()
    want = torch.relu(want)
    assert rel_err(to_np(outs[2][:, :N].float()), to_np(want)) < BF16_TOL
