"""Per-kernel parity of the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import np_ops
from oracle import x3d_oracle as O
from tests.gpu_util import (assert_close, bf16_round, dev, ncdhw64, rel_err, to_dev, to_np)

pytestmark = pytest.mark.gpu
DT = [torch.float32, torch.bfloat16]


def _ops():
    from x3d_tf_b200 import ops
    return ops


def _q(a, dtype):
    return bf16_round(a) if dtype == torch.bfloat16 else np.asarray(a, np.float32)


# ------------------------------------------------------------------------------- stem
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("N,T,H,W,C", [(2, 4, 32, 32, 24), (1, 5, 37, 45, 24), (1, 13, 18, 50, 32),
                                       (2, 1, 8, 8, 24), (1, 3, 33, 31, 40)])
def test_stem(dtype, N, T, H, W, C):
    rng = np.random.default_rng(N * 1000 + H)
    x = _q(rng.normal(size=(N, T, H, W, 3)), dtype)
    ks = rng.normal(size=(1, 3, 3, 3, C)).astype(np.float32) * 0.3
    kt = rng.normal(size=(5, 1, 1, 1, C)).astype(np.float32) * 0.5
    bias = rng.normal(size=C).astype(np.float32) * 0.2
    want = np.maximum(np_ops.stem_convs(x, ks, kt) + bias, 0.0)
    got = _ops().stem_fwd(to_dev(x, dtype), to_dev(ks.reshape(27, C)), to_dev(kt.reshape(5, C)),
                          to_dev(bias), dtype)
    assert_close(to_np(got), want, dtype, "stem")


def test_stem_mixed_dtypes():
    rng = np.random.default_rng(7)
    x = rng.normal(size=(1, 4, 20, 20, 3)).astype(np.float32)
    ks = rng.normal(size=(1, 3, 3, 3, 24)).astype(np.float32) * 0.3
    kt = rng.normal(size=(5, 1, 1, 1, 24)).astype(np.float32) * 0.5
    bias = np.zeros(24, np.float32)
    want = np.maximum(np_ops.stem_convs(x, ks, kt), 0.0)
    got = _ops().stem_fwd(to_dev(x), to_dev(ks.reshape(27, 24)), to_dev(kt.reshape(5, 24)),
                          to_dev(bias), torch.bfloat16)          # fp32 clips -> bf16 activations
    assert got.dtype == torch.bfloat16
    assert_close(to_np(got), want, torch.bfloat16, "stem f32->bf16")


def _pack_stem_tc(ks, kt, C):
    ws = ks.reshape(27, C).astype(np.float64)
    wt = kt.reshape(5, C).astype(np.float64)
    wc = np.zeros((5, 4, 32, 8), np.float64)
    for k in range(27):
        wc[:, k // 8, :C, k % 8] = ws[k][None, :] * wt
    return to_dev(wc, torch.bfloat16)


@pytest.mark.parametrize("in_dtype", DT)
@pytest.mark.parametrize("N,T,H,W,C", [(2, 4, 32, 32, 24), (1, 5, 37, 45, 24), (1, 13, 18, 50, 32),
                                       (2, 1, 8, 8, 24), (1, 16, 64, 64, 24), (1, 3, 91, 31, 24)])
def test_stem_tcgen05(in_dtype, N, T, H, W, C):
    rng = np.random.default_rng(N * 1000 + H + 1)
    x = _q(rng.normal(size=(N, T, H, W, 3)), in_dtype)
    ks = rng.normal(size=(1, 3, 3, 3, C)).astype(np.float32) * 0.3
    kt = rng.normal(size=(5, 1, 1, 1, C)).astype(np.float32) * 0.5
    bias = rng.normal(size=C).astype(np.float32) * 0.2
    want = np.maximum(np_ops.stem_convs(x, ks, kt) + bias, 0.0)
    got = _ops().stem_tc_fwd(to_dev(x, in_dtype), _pack_stem_tc(ks, kt, C), to_dev(bias))
    torch.cuda.synchronize()
    assert got.dtype == torch.bfloat16 and got.shape == want.shape
    # operands are rounded to bf16 (clip values and merged weights): error ~ 2^-8 of the scale
    assert rel_err(to_np(got), want) < 1.5e-2


# ------------------------------------------------------------------------------- channelwise
DW_CASES = [  # N, T, H, W, C, stride  -- covers both SAME-pad cases, odd extents, all strip widths
    (2, 4, 8, 8, 56, 1), (1, 16, 14, 14, 216, 1), (1, 13, 23, 23, 112, 1), (1, 4, 7, 7, 432, 1),
    (1, 5, 12, 6, 56, 1), (1, 4, 16, 16, 56, 2), (1, 13, 23, 23, 112, 2), (1, 4, 46, 46, 56, 2),
    (2, 3, 9, 13, 168, 2), (1, 1, 5, 5, 8, 1), (1, 2, 3, 20, 632, 1), (1, 16, 28, 28, 112, 1)]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("N,T,H,W,C,stride", DW_CASES)
def test_channelwise(dtype, N, T, H, W, C, stride):
    from x3d_tf_b200.arch import same_pad
    rng = np.random.default_rng(H * 100 + W + C + stride)
    x = _q(rng.normal(size=(N, T, H, W, C)), dtype)
    k = rng.normal(size=(3, 3, 3, 1, C)).astype(np.float32) * 0.3
    bias = rng.normal(size=C).astype(np.float32) * 0.2
    want = np_ops.channelwise_conv_same(x, k, stride) + bias
    _, ph, _ = same_pad(H, 3, stride)
    _, pw, _ = same_pad(W, 3, stride)
    out, partial = _ops().dw_fwd(to_dev(x, dtype), to_dev(k.reshape(27, C)), to_dev(bias), stride,
                                 ph, pw, True)
    assert_close(to_np(out), want, dtype, "channelwise")
    # SE partial sums: fp32 sums of the unrounded outputs
    sums = to_np(partial).astype(np.float64).sum(1)
    np.testing.assert_allclose(sums, want.sum((1, 2, 3)), rtol=2e-4,
                               atol=2e-4 * np.abs(want).sum((1, 2, 3)).max())
    out2, p2 = _ops().dw_fwd(to_dev(x, dtype), to_dev(k.reshape(27, C)), to_dev(bias), stride,
                             ph, pw, False)
    assert p2 is None and torch.equal(out, out2)
    # swish fused into the epilogue (blocks without SE, model.py:316): swish of the unrounded output
    out3, _ = _ops().dw_fwd(to_dev(x, dtype), to_dev(k.reshape(27, C)), to_dev(bias), stride,
                            ph, pw, False, swish=True)
    assert_close(to_np(out3), want / (1.0 + np.exp(-want)), dtype, "channelwise + swish")
    with pytest.raises(ValueError):
        _ops().dw_fwd(to_dev(x, dtype), to_dev(k.reshape(27, C)), to_dev(bias), stride, ph, pw, True, swish=True)


@pytest.mark.parametrize("N,T,H,W,C,stride", DW_CASES + [(3, 4, 64, 64, 56, 1), (2, 16, 33, 70, 24, 1), (2, 5, 64, 64, 56, 2),
                                                        (5, 3, 8, 8, 432, 1), (2, 7, 16, 16, 216, 2), (2, 4, 56, 56, 56, 1),
                                                        (3, 5, 28, 28, 112, 1), (3, 6, 14, 14, 216, 1),   # 7 rows per thread
                                                        (4, 6, 7, 7, 216, 1), (9, 4, 8, 6, 56, 1), (1, 3, 5, 8, 24, 1)])   # four clips per item
def test_channelwise_planar(N, T, H, W, C, stride):
    """x3d_dw3x3x3_planar_fwd (lanes = pixels, taps in uniform registers) against the same oracle as
    x3d_dw3x3x3_fwd: output, SE partial sums, swish epilogue; several clips / tiles / chunks per CTA."""
    from x3d_tf_b200.arch import same_pad
    ops = _ops()
    if ops.dw_planar_supported(T, H, W, C, stride) <= 0:
        pytest.skip("planar kernel does not take this width (C > 576)")
    dtype = torch.bfloat16
    rng = np.random.default_rng(H * 100 + W + C + stride)
    x = _q(rng.normal(size=(N, T, H, W, C)), dtype)
    k = rng.normal(size=(3, 3, 3, 1, C)).astype(np.float32) * 0.3
    bias = rng.normal(size=C).astype(np.float32) * 0.2
    want = np_ops.channelwise_conv_same(x, k, stride) + bias
    _, ph, _ = same_pad(H, 3, stride)
    _, pw, _ = same_pad(W, 3, stride)
    taps = ops.dw_planar_taps(to_dev(k.reshape(27, C)), to_dev(bias))
    out, partial = ops.dw_planar_fwd(to_dev(x, dtype), taps, stride, ph, pw, True)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, dtype, "planar channelwise")
    sums = to_np(partial).astype(np.float64).sum(1)
    np.testing.assert_allclose(sums, want.sum((1, 2, 3)), rtol=2e-4,
                               atol=2e-4 * np.abs(want).sum((1, 2, 3)).max())
    out2, p2 = ops.dw_planar_fwd(to_dev(x, dtype), taps, stride, ph, pw, False)
    assert p2 is None and torch.equal(out, out2)
    out3, _ = ops.dw_planar_fwd(to_dev(x, dtype), taps, stride, ph, pw, False, swish=True)
    assert_close(to_np(out3), want / (1.0 + np.exp(-want)), dtype, "planar channelwise + swish")


# ------------------------------------------------------------------------------- fused expand + channelwise
AB_CASES = [  # N, T, H, W, Cin, C, stride  -- every stage shape class, odd extents, both pads, K > 64
    (2, 4, 16, 16, 24, 56, 1), (1, 16, 14, 14, 96, 216, 1), (1, 13, 23, 23, 48, 112, 1),
    (1, 4, 8, 8, 192, 432, 1), (1, 5, 12, 9, 24, 56, 1), (1, 4, 32, 32, 24, 56, 2),
    (1, 13, 23, 23, 24, 112, 2), (1, 4, 46, 46, 24, 56, 2), (2, 3, 18, 26, 96, 432, 2),
    (1, 1, 16, 16, 48, 216, 1), (1, 2, 9, 20, 136, 632, 1), (1, 16, 28, 28, 48, 112, 1),
    (1, 3, 20, 20, 72, 168, 2)]


@pytest.mark.parametrize("N,T,H,W,Cin,C,stride", AB_CASES)
def test_fused_expand_channelwise_tcgen05(N, T, H, W, Cin, C, stride):
    """x3d_expand_dw_fwd == bf16(relu(x.Wa + ta)) -> channelwise conv, the unfused pair's result."""
    from x3d_tf_b200.arch import same_pad
    ops = _ops()
    if ops.expand_dw_supported(T, H, W, Cin, C, stride) <= 0:
        pytest.skip("no fused tile plan for this shape (the model falls back to the two kernels)")
    rng = np.random.default_rng(H * 100 + W + C + stride + Cin)
    x = bf16_round(rng.normal(size=(N, T, H, W, Cin)))
    wa = bf16_round(rng.normal(size=(Cin, C)) / np.sqrt(Cin))
    ta = rng.normal(size=C).astype(np.float32) * 0.3
    k = rng.normal(size=(3, 3, 3, 1, C)).astype(np.float32) * 0.3
    tb = rng.normal(size=C).astype(np.float32) * 0.2
    a = np.maximum(x.reshape(-1, Cin).astype(np.float64) @ wa.astype(np.float64) + ta, 0.0)
    a = bf16_round(a).reshape(N, T, H, W, C)                     # the ring holds bf16, like HBM would
    want = np_ops.channelwise_conv_same(a, k, stride) + tb
    npad, kpad = (C + 15) // 16 * 16, (Cin + 63) // 64 * 64
    wp = np.zeros((npad, kpad), np.float32)
    wp[:C, :Cin] = wa.T
    _, ph, _ = same_pad(H, 3, stride)
    _, pw, _ = same_pad(W, 3, stride)
    out, partial = ops.expand_dw_fwd(to_dev(x, torch.bfloat16), to_dev(wp, torch.bfloat16), to_dev(ta),
                                     to_dev(k.reshape(27, C)), to_dev(tb), stride, ph, pw, True)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, torch.bfloat16, "fused expand+channelwise")
    sums = to_np(partial).astype(np.float64).sum(1)
    np.testing.assert_allclose(sums, want.sum((1, 2, 3)), rtol=2e-4,
                               atol=2e-4 * np.abs(want).sum((1, 2, 3)).max())
    out2, p2 = ops.expand_dw_fwd(to_dev(x, torch.bfloat16), to_dev(wp, torch.bfloat16), to_dev(ta),
                                 to_dev(k.reshape(27, C)), to_dev(tb), stride, ph, pw, False)
    assert p2 is None and torch.equal(out, out2)


@pytest.mark.parametrize("N,T,H,W,Cin,C,stride", AB_CASES + [(3, 4, 20, 20, 24, 56, 1), (5, 2, 16, 16, 48, 112, 2),
                                                            (2, 16, 64, 64, 24, 56, 1), (2, 7, 64, 64, 24, 56, 2)])
def test_fused_expand_channelwise_persistent(N, T, H, W, Cin, C, stride):
    """x3d_expand_dw2_fwd == relu(x.Wa + ta) (kept in fp32, NOT rounded to bf16) -> channelwise
    conv (+ bias, + swish), i.e. model.py:306-316 with one rounding at the output; several clips so
    that a CTA walks more than one work item."""
    from x3d_tf_b200.arch import same_pad
    ops = _ops()
    if ops.expand_dw2_supported(T, H, W, Cin, C, stride) <= 0:
        pytest.skip("no fused tile plan for this shape (the model falls back to the two kernels)")
    rng = np.random.default_rng(H * 100 + W + C + stride + Cin)
    x = bf16_round(rng.normal(size=(N, T, H, W, Cin)))
    wa = bf16_round(rng.normal(size=(Cin, C)) / np.sqrt(Cin))
    ta = rng.normal(size=C).astype(np.float32) * 0.3
    k = rng.normal(size=(3, 3, 3, 1, C)).astype(np.float32) * 0.3
    tb = rng.normal(size=C).astype(np.float32) * 0.2
    a = np.maximum(x.reshape(-1, Cin).astype(np.float64) @ wa.astype(np.float64) + ta, 0.0)
    want = np_ops.channelwise_conv_same(a.reshape(N, T, H, W, C), k, stride) + tb
    npad, kpad = (C + 15) // 16 * 16, (Cin + 63) // 64 * 64
    wp = np.zeros((npad, kpad), np.float32)
    wp[:C, :Cin] = wa.T
    _, ph, _ = same_pad(H, 3, stride)
    _, pw, _ = same_pad(W, 3, stride)
    args = (to_dev(x, torch.bfloat16), to_dev(wp, torch.bfloat16), to_dev(ta), to_dev(k.reshape(27, C)),
            to_dev(tb), stride, ph, pw)
    out, partial = ops.expand_dw2_fwd(*args, True)
    torch.cuda.synchronize()
    assert_close(to_np(out), want, torch.bfloat16, "persistent fused expand+channelwise")
    sums = to_np(partial).astype(np.float64).sum(1)
    np.testing.assert_allclose(sums, want.sum((1, 2, 3)), rtol=2e-4,
                               atol=2e-4 * np.abs(want).sum((1, 2, 3)).max())
    out2, p2 = ops.expand_dw2_fwd(*args, False)
    assert p2 is None and torch.equal(out, out2)
    out3, _ = ops.expand_dw2_fwd(*args, False, swish=True)
    assert_close(to_np(out3), want / (1.0 + np.exp(-want)), torch.bfloat16, "persistent fused + swish")
    with pytest.raises(ValueError):
        ops.expand_dw2_fwd(*args, True, swish=True)


# ------------------------------------------------------------------------------- pointwise
def _pw_ref(a, w, bias, res=None, se=None, rpc=0, swish=False, relu=False):
    a = np.asarray(a, np.float64)
    if se is not None:
        a = a * np.repeat(se.astype(np.float64), rpc, axis=0)[: a.shape[0]]
    if swish:
        a = a / (1.0 + np.exp(-a))
    y = a @ w.astype(np.float64) + bias
    if res is not None:
        y = y + res
    return np.maximum(y, 0) if relu else y


PW_CASES = [(1000, 24, 56), (300, 56, 24), (4097, 112, 48), (129, 216, 96), (784, 192, 432),
            (50, 432, 192), (128, 8, 8), (77, 632, 280)]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,K,N", PW_CASES)
def test_pointwise_simt_plain(dtype, M, K, N):
    rng = np.random.default_rng(M + K + N)
    a = _q(rng.normal(size=(M, K)), dtype)
    w = rng.normal(size=(K, N)).astype(np.float32) / np.sqrt(K)
    bias = rng.normal(size=N).astype(np.float32)
    got = _ops().pw_fwd(to_dev(a, dtype), to_dev(w), to_dev(bias), M=M, K=K, Nc=N, relu=True)
    assert_close(to_np(got), _pw_ref(a, w, bias, relu=True), dtype, "pw simt")


@pytest.mark.parametrize("dtype", DT)
def test_pointwise_simt_prologue_epilogue(dtype):
    rng = np.random.default_rng(5)
    M, K, N, rpc = 600, 112, 48, 200
    a = _q(rng.normal(size=(M, K)), dtype)
    w = rng.normal(size=(K, N)).astype(np.float32) / np.sqrt(K)
    bias = rng.normal(size=N).astype(np.float32)
    res = _q(rng.normal(size=(M, N)), dtype)
    se = rng.uniform(0.1, 0.9, size=(3, K)).astype(np.float32)
    got = _ops().pw_fwd(to_dev(a, dtype), to_dev(w), to_dev(bias), M=M, K=K, Nc=N,
                        residual=to_dev(res, dtype), se=to_dev(se), rows_per_clip=rpc, swish=True,
                        relu=True)
    assert_close(to_np(got), _pw_ref(a, w, bias, res, se, rpc, True, True), dtype, "pw simt pro/epi")
    got = _ops().pw_fwd(to_dev(a, dtype), to_dev(w), to_dev(bias), M=M, K=K, Nc=N, swish=True)
    assert_close(to_np(got), _pw_ref(a, w, bias, swish=True), dtype, "pw simt swish only")


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("H,W", [(8, 8), (7, 9), (23, 23)])
def test_pointwise_strided_gather(dtype, H, W):
    rng = np.random.default_rng(H + W)
    N, T, K, Nc = 2, 3, 24, 48
    x = _q(rng.normal(size=(N, T, H, W, K)), dtype)
    k = rng.normal(size=(1, 1, 1, K, Nc)).astype(np.float32) / 5
    bias = rng.normal(size=Nc).astype(np.float32)
    want = np_ops.pointwise_conv_valid(x, k, 2) + bias
    Ho, Wo = want.shape[2:4]
    got = _ops().pw_fwd(to_dev(x, dtype), to_dev(k.reshape(K, Nc)), to_dev(bias),
                        M=N * T * Ho * Wo, K=K, Nc=Nc, gather=(T, Ho, Wo, H, W, 2))
    assert_close(to_np(got).reshape(want.shape), want, dtype, "shortcut gather")


def _pack_tc(w, scale=None):
    K, N = w.shape
    npad, kpad = (N + 15) // 16 * 16, (K + 63) // 64 * 64
    wp = np.zeros((npad, kpad), np.float32)
    wp[:N, :K] = w.T
    return to_dev(wp, torch.bfloat16)


@pytest.mark.parametrize("M,K,N", PW_CASES + [(128 * 300 + 5, 56, 24), (20000, 24, 56)])
def test_pointwise_tcgen05_plain(M, K, N):
    rng = np.random.default_rng(M + K + N)
    a = bf16_round(rng.normal(size=(M, K)))
    w = bf16_round(rng.normal(size=(K, N)) / np.sqrt(K))
    bias = rng.normal(size=N).astype(np.float32)
    got = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), M=M, K=K, Nc=N,
                           relu=True)
    torch.cuda.synchronize()
    assert_close(to_np(got), _pw_ref(a, w, bias, relu=True), torch.bfloat16, "pw tcgen05")


def test_pointwise_tcgen05_prologue_epilogue():
    rng = np.random.default_rng(11)
    for (M, K, N, rpc) in [(600, 112, 48, 200), (5000, 216, 96, 784), (1500, 56, 24, 100),
                           (900, 432, 192, 300)]:
        a = bf16_round(rng.normal(size=(M, K)))
        w = bf16_round(rng.normal(size=(K, N)) / np.sqrt(K))
        bias = rng.normal(size=N).astype(np.float32)
        res = bf16_round(rng.normal(size=(M, N)))
        nclip = -(-M // rpc)
        se = rng.uniform(0.1, 0.9, size=(nclip, K)).astype(np.float32)
        got = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), M=M, K=K,
                               Nc=N, residual=to_dev(res, torch.bfloat16), se=to_dev(se),
                               rows_per_clip=rpc, swish=True, relu=True)
        torch.cuda.synchronize()
        # the prologue rounds swish(se*a) to bf16 before the MMA: allow one more bf16 ulp of scale
        want = _pw_ref(a, w, bias, res, se, rpc, True, True)
        assert rel_err(to_np(got), want) < 1.2e-2, (M, K, N)
        got2 = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), M=M, K=K,
                                Nc=N, swish=True)
        assert rel_err(to_np(got2), _pw_ref(a, w, bias, swish=True)) < 1.2e-2


@pytest.mark.parametrize("NT,Hi,Wi,K,K2,N,stride,pro", [
    (6, 32, 32, 112, 24, 48, 2, False),      # 16x16 frames: half a frame per tile
    (8, 16, 16, 216, 48, 96, 2, True),       # 8x8 frames: two frames per tile
    (5, 16, 16, 432, 96, 192, 2, False),     # odd frame count: last tile half outside the tensor
    (3, 127, 128, 56, 24, 24, 2, True),      # odd height: Ho = 64, last sampled row is 126
    (2, 64, 64, 56, 24, 24, 1, True),        # stride 1 (channel change only)
    (3, 30, 44, 112, 24, 48, 2, True),       # 15x22 frames: no aligned tiles -> gathered rows as a dense second source
])
def test_pointwise_tcgen05_shortcut_as_extra_k(NT, Hi, Wi, K, K2, N, stride, pro):
    """x3d_pw_tc_fwd's second source: relu(bias + pro(A).W + sample(X).W2) against the fp64 formula."""
    rng = np.random.default_rng(NT * Hi + K)
    ops = _ops()
    sampled = ops.pw_tc_sampler_supported(Hi, Wi, stride)
    assert sampled == (Wi != 44)
    Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
    M = NT * Ho * Wo
    a = bf16_round(rng.normal(size=(M, K)))
    x = bf16_round(rng.normal(size=(1, NT, Hi, Wi, K2)))
    w = bf16_round(rng.normal(size=(K, N)) / np.sqrt(K))
    w2 = bf16_round(rng.normal(size=(K2, N)) / np.sqrt(K2))
    bias = rng.normal(size=N).astype(np.float32)
    k1 = (K + 63) // 64 * 64
    wp = np.zeros(((N + 15) // 16 * 16, k1 + (K2 + 63) // 64 * 64), np.float32)
    wp[:N, :K] = w.T
    wp[:N, k1:k1 + K2] = w2.T
    rpc = 2 * Ho * Wo                                           # two frames per "clip"
    se = rng.uniform(0.1, 0.9, size=(-(-M // rpc), K)).astype(np.float32) if pro else None
    got = ops.pw_tc_fwd(to_dev(a, torch.bfloat16), to_dev(wp, torch.bfloat16), to_dev(bias), M=M, K=K, Nc=N,
                        se=to_dev(se) if pro else None, rows_per_clip=rpc if pro else 0, swish=pro, relu=True,
                        a2=to_dev(x, torch.bfloat16) if sampled else ops.gather_rows_fwd(to_dev(x, torch.bfloat16), stride),
                        a2_stride=stride)
    torch.cuda.synchronize()
    xs = x[0, :, ::stride, ::stride, :].reshape(M, K2).astype(np.float64)
    want = _pw_ref(a, w, bias, xs @ w2.astype(np.float64), se, rpc, pro, True)
    assert rel_err(to_np(got), want) < (1.2e-2 if pro else 2.0 ** -7), (NT, Hi, Wi)
    assert not ops.pw_tc_sampler_supported(112, 112, 2) and not ops.pw_tc_sampler_supported(91, 91, 2)


@pytest.mark.parametrize("M,K,N", [(1024, 192, 432), (1000, 96, 56), (130, 24, 280)])
def test_pointwise_tcgen05_column_means(M, K, N):
    """conv_5 + pool_5 epilogue: means over 64-row groups of the bf16-rounded output (rows >= M count 0)."""
    rng = np.random.default_rng(M + N)
    a = bf16_round(rng.normal(size=(M, K)))
    w = bf16_round(rng.normal(size=(K, N)) / np.sqrt(K))
    bias = rng.normal(size=N).astype(np.float32)
    args = dict(M=M, K=K, Nc=N, relu=True, colmean=True)
    out, means = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), **args)
    only = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), store=False, **args)
    torch.cuda.synchronize()
    assert_close(to_np(out), _pw_ref(a, w, bias, relu=True), torch.bfloat16, "pw tcgen05 + means")
    groups = -(-M // 64)
    padded = np.zeros((groups * 64, N))
    padded[:M] = to_np(out).astype(np.float64)
    want = padded.reshape(groups, 64, N).sum(1) / 64
    np.testing.assert_allclose(to_np(means)[:groups], want, rtol=2e-6, atol=1e-6)
    assert torch.equal(means[:groups], only[:groups])


def test_pointwise_tcgen05_matches_simt_bitwise_inputs():
    """Same bf16 inputs through both pointwise kernels: results agree to bf16 rounding."""
    rng = np.random.default_rng(3)
    M, K, N = 3000, 96, 216
    a = bf16_round(rng.normal(size=(M, K)))
    w = bf16_round(rng.normal(size=(K, N)) / np.sqrt(K))
    bias = rng.normal(size=N).astype(np.float32)
    t1 = _ops().pw_tc_fwd(to_dev(a, torch.bfloat16), _pack_tc(w), to_dev(bias), M=M, K=K, Nc=N)
    t2 = _ops().pw_fwd(to_dev(a, torch.bfloat16), to_dev(w), to_dev(bias), M=M, K=K, Nc=N)
    assert rel_err(to_np(t1), to_np(t2).astype(np.float64)) < 2.0 ** -7


# ------------------------------------------------------------------------------- small kernels
def test_se_mlp():
    rng = np.random.default_rng(2)
    N, nblk, C, Cw = 3, 5, 112, 8
    partial = rng.normal(size=(N, nblk, C)).astype(np.float32)
    w1 = rng.normal(size=(C, Cw)).astype(np.float32) / 5
    b1 = rng.normal(size=Cw).astype(np.float32)
    w2 = rng.normal(size=(Cw, C)).astype(np.float32)
    b2 = rng.normal(size=C).astype(np.float32)
    count = 77
    mean = partial.astype(np.float64).sum(1) / count
    z = np.maximum(mean @ w1 + b1, 0)
    want = 1 / (1 + np.exp(-(z @ w2 + b2)))
    got = _ops().se_mlp_fwd(to_dev(partial), count, to_dev(w1), to_dev(b1), to_dev(w2), to_dev(b2))
    np.testing.assert_allclose(to_np(got), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", DT)
def test_avgpool(dtype):
    rng = np.random.default_rng(4)
    x = _q(rng.normal(size=(3, 4, 7, 7, 432)), dtype)
    got = _ops().avgpool_fwd(to_dev(x, dtype))
    np.testing.assert_allclose(to_np(got), x.astype(np.float64).mean((1, 2, 3)), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("num_preds", [1, 2, 10])
def test_softmax_viewmean(num_preds):
    rng = np.random.default_rng(num_preds)
    lg = (rng.normal(size=(20, 400)) * 4).astype(np.float32)
    p = np.exp(lg.astype(np.float64) - lg.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    want = p.reshape(-1, num_preds, 400).mean(1)
    got = _ops().softmax_viewmean_fwd(to_dev(lg), num_preds)
    np.testing.assert_allclose(to_np(got), want, rtol=2e-6, atol=1e-9)
    if num_preds == 2:
        with pytest.raises(ValueError):
            _ops().softmax_viewmean_fwd(to_dev(lg[:19]), 2)


def test_error_reporting():
    from x3d_tf_b200._lib import X3DLibError
    x = torch.zeros((1, 2, 4, 4, 12), device=dev())            # C not a multiple of 8
    with pytest.raises(X3DLibError, match="multiple of 8"):
        _ops().dw_fwd(x, torch.zeros((27, 12), device=dev()), torch.zeros(12, device=dev()), 1, 1, 1, False)
    with pytest.raises(ValueError):
        _ops().avgpool_fwd(torch.zeros((1, 2, 2, 2, 8)))       # CPU tensor: no CPU path


# ------------------------------------------------------------------------------- gather / head
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("H,W,stride,C", [(8, 8, 2, 24), (7, 9, 2, 48), (23, 23, 2, 96), (5, 6, 1, 24),
                                          (91, 91, 2, 24)])
def test_gather_rows_is_bit_exact(dtype, H, W, stride, C):
    """Index work: the sampled rows must be copied bit for bit (model.py:360-367, 'valid',
    stride (1,s,s) => positions 0, s, 2s, ...)."""
    rng = np.random.default_rng(H * W + C)
    x = to_dev(rng.normal(size=(2, 3, H, W, C)), dtype)
    got = _ops().gather_rows_fwd(x, stride)
    want = x[:, :, ::stride, ::stride, :].contiguous().view(-1, C)
    assert got.shape == want.shape
    assert torch.equal(got, want)


@pytest.mark.parametrize("M,K,N,relu", [(80, 432, 2048, True), (80, 2048, 400, False), (1, 432, 2048, True),
                                        (7, 64, 20, False), (130, 72, 36, True), (300, 2200, 52, False)])
def test_head_fc(M, K, N, relu):
    rng = np.random.default_rng(M + K + N)
    a = rng.normal(size=(M, K)).astype(np.float32)
    w = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    want = a.astype(np.float64) @ w.astype(np.float64) + b
    if relu:
        want = np.maximum(want, 0)
    got = _ops().head_fc_fwd(to_dev(a), to_dev(w), to_dev(b), K=K, Nc=N, relu=relu)
    assert_close(to_np(got), want, torch.float32, "head fc")


def test_shortcut_tc_path_matches_oracle():
    """ResBlock shortcut (residual + bn_r, model.py:386-388) through gather + tcgen05 GEMM."""
    rng = np.random.default_rng(5)
    N, T, H, W, K, Nc = 2, 3, 23, 17, 24, 48
    x = bf16_round(rng.normal(size=(N, T, H, W, K)))
    k = rng.normal(size=(1, 1, 1, K, Nc)).astype(np.float32) / 5
    bias = rng.normal(size=Nc).astype(np.float32)
    want = np_ops.pointwise_conv_valid(x, bf16_round(k), 2) + bias
    rows = _ops().gather_rows_fwd(to_dev(x, torch.bfloat16), 2)
    got = _ops().pw_tc_fwd(rows, _pack_tc(k.reshape(K, Nc)), to_dev(bias), M=rows.shape[0], K=K, Nc=Nc)
    assert_close(to_np(got).reshape(want.shape), want, torch.bfloat16, "shortcut tc")
