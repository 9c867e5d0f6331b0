"""CPU ORACLE for the X3D forward path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module.  The product package (`x3d_tf_b200`) never does.

What it is: a restatement, in torch-CPU float64/float32, of the arithmetic of the reference
`model.py` (TensorFlow 2.4.1 Keras).  TensorFlow is an un-vendored dependency
(`requirements.txt:4`, tensorflow==2.4.1) that is absent here and cannot be installed, so the
layer semantics below restate TF's published behaviour (Conv3D `valid`/`same`, grouped conv,
BatchNormalization inference, GlobalAveragePooling3D, Dense, Softmax) at the reference's own
call sites, each cited as `model.py:line`.

PARITY PARTIALLY PINNED.  The reference ships no tests and no numeric golden vectors for this
path; TensorFlow and the checkpoint data shards are absent, so no output
of the real Keras/TensorFlow stack can be generated here.  What IS pinned:
  * `tests/golden/ref_*.npz`: outputs of the reference's OWN, unmodified `model.py` executed in the
    build container on `oracle/tf_shim` (numpy float64 Keras primitives), for all five variants --
    everything `model.py`/`utils.py`/`configs/default.py` decide (graph, rounding, SE placement by
    the global block counter, call order, view averaging, variable names) is the reference's code;
    this oracle matches those logits to 1e-10 (`tests/test_reference_golden.py`);
  * the structural golden data the reference ships (per-stage shapes and parameter counts in
    `models/*/X3D_*.txt`, every variable name/shape of `models/*/model.index`);
  * a second, independent numpy implementation (`oracle/np_ops.py`) of the two layers whose TF
    semantics are not the torch default (SAME-padded strided channelwise conv; the padded stem).
Still unpinned: the numerics of TensorFlow's own kernels, restated from documented TF semantics.

Weights are a dict {checkpoint attribute path -> numpy float32 array}, layouts as TF stores
them (conv kernels DHWIO, dense [in,out]); names per SURVEY.md Appendix C.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# utils.py:7-40
def round_width(width, multiplier, min_depth=8, divisor=8):
    if not multiplier:
        return width
    width *= multiplier
    min_depth = min_depth or divisor
    new_filters = max(min_depth, int(width + divisor / 2) // divisor * divisor)
    if new_filters < 0.9 * width:
        new_filters += divisor
    return int(new_filters)


def round_repeats(repeats, multiplier):
    if not multiplier:
        return repeats
    return int(math.ceil(multiplier * repeats))


# --------------------------------------------------------------------------------------
class OracleSpec:
    """Graph hyper-parameters as `X3D.__init__` derives them (model.py:21-76)."""

    def __init__(self, *, width_factor, depth_factor, bottleneck_factor, c1_channels=12,
                 scale_res2=False, num_classes=400, bn_eps=1e-5, temp_filter=5,
                 num_temporal_views=1, num_spatial_crops=1, se_ratio=0.0625,
                 first_block_index=1):
        self.num_classes = num_classes
        self.bn_eps = bn_eps
        self.temp_filter = temp_filter
        self.num_preds = num_temporal_views * num_spatial_crops          # model.py:25
        if scale_res2:                                                    # model.py:31-37
            self.conv1_dim = round_width(c1_channels, width_factor)
            mult = 1
        else:
            self.conv1_dim = round_width(c1_channels, 2)
            mult = 2
        b = c1_channels * mult
        basis = [[1, b], [2, round_width(b, 2)], [5, round_width(b, 4)],
                 [3, round_width(b, 8)]]                                  # model.py:40-44
        self.stages: List[Tuple[int, int, int, int]] = []                 # depth,in,inner,out
        out_dim = self.conv1_dim
        for depth_b, ch_b in basis:                                       # model.py:59-76
            in_dim = out_dim
            out_dim = round_width(ch_b, width_factor)
            inner = int(out_dim * bottleneck_factor)
            depth = round_repeats(depth_b, depth_factor)
            self.stages.append((depth, in_dim, inner, out_dim))
        # global block counter: model.py:326 (starts 0), :351 (incremented before use), :378
        self.blocks = []      # (stage, j, cin, inner, cout, stride, se_width or 0, shortcut)
        g = first_block_index
        for s, (depth, in_dim, inner, out_dim) in enumerate(self.stages):
            for j in range(depth):
                cin = in_dim if j == 0 else out_dim                       # model.py:448-452
                stride = 2 if j == 0 else 1
                se = round_width(inner, se_ratio) if (g + 1) % 2 == 0 else 0   # model.py:275-276
                shortcut = cin != out_dim or stride != 1                  # model.py:359
                self.blocks.append((s, j, cin, inner, out_dim, stride, se, shortcut))
                g += 1
        self.conv5_dim = self.stages[-1][2]                               # model.py:81

    @classmethod
    def from_cfg(cls, cfg, **kw):
        n = cfg["NETWORK"] if isinstance(cfg, dict) else cfg.NETWORK
        t = cfg["TEST"] if isinstance(cfg, dict) else cfg.TEST
        g = (lambda node, k: node[k])
        return cls(width_factor=g(n, "WIDTH_FACTOR"), depth_factor=g(n, "DEPTH_FACTOR"),
                   bottleneck_factor=g(n, "BOTTLENECK_WIDTH_FACTOR"),
                   c1_channels=g(n, "C1_CHANNELS"), scale_res2=g(n, "SCALE_RES2"),
                   num_classes=g(n, "NUM_CLASSES"), bn_eps=g(n["BN"], "EPS"),
                   temp_filter=g(n, "C1_TEMP_FILTER"),
                   num_temporal_views=g(t, "NUM_TEMPORAL_VIEWS"),
                   num_spatial_crops=g(t, "NUM_SPATIAL_CROPS"), **kw)

    def variable_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """All checkpoint variables."""
        v: Dict[str, Tuple[int, ...]] = {}

        def bn(p, c):
            for n in ("gamma", "beta", "moving_mean", "moving_variance"):
                v[f"{p}/{n}"] = (c,)
        c1 = self.conv1_dim
        v["conv1/conv_s/kernel"] = (1, 3, 3, 3, c1)                       # model.py:178-184
        v["conv1/conv_t/kernel"] = (self.temp_filter, 1, 1, 1, c1)        # model.py:187-194
        bn("conv1/bn", c1)
        for (s, j, cin, inner, cout, stride, se, shortcut) in self.blocks:
            p = f"stages/{s}/stage/layer_with_weights-{j}"
            if shortcut:
                v[f"{p}/residual/kernel"] = (1, 1, 1, cin, cout)          # model.py:360-367
                bn(f"{p}/bn_r", cout)
            q = p + "/bottleneck"
            v[f"{q}/a/kernel"] = (1, 1, 1, cin, inner)                    # model.py:246-253
            bn(f"{q}/bn_a", inner)
            v[f"{q}/b/kernel"] = (3, 3, 3, 1, inner)                      # model.py:259-267
            bn(f"{q}/bn_b", inner)
            if se:
                v[f"{q}/se_fc1/kernel"] = (1, 1, 1, inner, se)            # model.py:278-283
                v[f"{q}/se_fc1/bias"] = (se,)
                v[f"{q}/se_fc2/kernel"] = (1, 1, 1, se, inner)            # model.py:284-290
                v[f"{q}/se_fc2/bias"] = (inner,)
            v[f"{q}/c/kernel"] = (1, 1, 1, inner, cout)                   # model.py:292-299
            bn(f"{q}/bn_c", cout)
        v["conv5/layer_with_weights-0/kernel"] = (1, 1, 1, self.stages[-1][3], self.conv5_dim)
        bn("conv5/layer_with_weights-1", self.conv5_dim)
        v["fc1/kernel"] = (1, 1, 1, self.conv5_dim, 2048)                 # model.py:95-102
        v["fc2/kernel"] = (2048, self.num_classes)                        # model.py:104-108
        v["fc2/bias"] = (self.num_classes,)
        return v


# --------------------------------------------------------------------------------------
# TF layer semantics on NCDHW torch tensors
def _k(w: np.ndarray, dtype) -> torch.Tensor:
    """DHWIO (TF) -> OIDHW (torch)."""
    return torch.from_numpy(np.ascontiguousarray(w)).to(dtype).permute(4, 3, 0, 1, 2).contiguous()


def _v(w: np.ndarray, dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w)).to(dtype)


def tf_same_pads(size: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME': out=ceil(in/s); total=max((out-1)*s+k-in,0); before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv3d_same(x, w, stride, groups):
    """Keras Conv3D(padding='same') -- model.py:246-253, 259-267, 292-299."""
    pads = []
    for dim, k, s in zip((4, 3, 2), (w.shape[4], w.shape[3], w.shape[2]),
                         (stride[2], stride[1], stride[0])):
        b, a = tf_same_pads(x.shape[dim], k, s)
        pads += [b, a]
    return F.conv3d(F.pad(x, pads), w, None, stride=stride, groups=groups)


def conv3d_valid(x, w, stride, groups=1):
    """Keras Conv3D(padding='valid') -- model.py:178-194, 360-367, conv5/fc1."""
    return F.conv3d(x, w, None, stride=stride, groups=groups)


def batchnorm_inference(x, W, prefix, eps, dtype):
    """Keras BatchNormalization(axis=-1), training=False: (x-mean)*gamma/sqrt(var+eps)+beta."""
    g, b = _v(W[prefix + "/gamma"], dtype), _v(W[prefix + "/beta"], dtype)
    m, v = _v(W[prefix + "/moving_mean"], dtype), _v(W[prefix + "/moving_variance"], dtype)
    sh = (1, -1, 1, 1, 1)
    return (x - m.view(sh)) * (g.view(sh) * torch.rsqrt(v.view(sh) + eps)) + b.view(sh)


def global_avg_pool(x):
    """AdaptiveAvgPool3D / GlobalAveragePooling3D -- model.py:473-483."""
    return x.mean(dim=(2, 3, 4), keepdim=True)


# --------------------------------------------------------------------------------------
def stem(W, x, spec: OracleSpec, dtype):
    """X3D_Stem.call -- model.py:202-210."""
    out = F.pad(x, (1, 1, 1, 1, 0, 0))                                    # model.py:203 (H,W by 1)
    out = conv3d_valid(out, _k(W["conv1/conv_s/kernel"], dtype), (1, 2, 2))        # :204
    tp = spec.temp_filter // 2
    out = F.pad(out, (0, 0, 0, 0, tp, tp))                                # :205
    out = conv3d_valid(out, _k(W["conv1/conv_t/kernel"], dtype), (1, 1, 1),
                       groups=spec.conv1_dim)                             # :206
    out = batchnorm_inference(out, W, "conv1/bn", spec.bn_eps, dtype)     # :207
    return F.relu(out)                                                    # :208


def bottleneck(W, x, p, inner, stride, se, spec, dtype, taps=None):
    """Bottleneck.call -- model.py:305-320."""
    q = p + "/bottleneck"
    out = conv3d_same(x, _k(W[q + "/a/kernel"], dtype), (1, 1, 1), 1)     # :306
    out = F.relu(batchnorm_inference(out, W, q + "/bn_a", spec.bn_eps, dtype))    # :307-308
    if taps is not None:
        taps[q + "/a_out"] = out
    out = conv3d_same(out, _k(W[q + "/b/kernel"], dtype), (1, stride, stride), inner)  # :309
    out = batchnorm_inference(out, W, q + "/bn_b", spec.bn_eps, dtype)    # :310
    if taps is not None:
        taps[q + "/b_out"] = out
    if se:                                                                # :311-315
        m = global_avg_pool(out)
        z = F.relu(F.conv3d(m, _k(W[q + "/se_fc1/kernel"], dtype), _v(W[q + "/se_fc1/bias"], dtype)))
        sc = torch.sigmoid(F.conv3d(z, _k(W[q + "/se_fc2/kernel"], dtype),
                                    _v(W[q + "/se_fc2/bias"], dtype)))
        if taps is not None:
            taps[q + "/se_scale"] = sc
        out = out * sc
    out = out * torch.sigmoid(out)                                        # swish, :316
    out = conv3d_same(out, _k(W[q + "/c/kernel"], dtype), (1, 1, 1), 1)   # :317
    return batchnorm_inference(out, W, q + "/bn_c", spec.bn_eps, dtype)   # :318


def res_block(W, x, blk, spec, dtype, taps=None):
    """ResBlock.call -- model.py:384-394."""
    (s, j, cin, inner, cout, stride, se, shortcut) = blk
    p = f"stages/{s}/stage/layer_with_weights-{j}"
    out = bottleneck(W, x, p, inner, stride, se, spec, dtype, taps)
    if shortcut:                                                          # :386-389
        res = conv3d_valid(x, _k(W[p + "/residual/kernel"], dtype), (1, stride, stride))
        res = batchnorm_inference(res, W, p + "/bn_r", spec.bn_eps, dtype)
        out = res + out
    else:
        out = x + out                                                     # :391
    return F.relu(out)                                                    # :392


def forward(W: Dict[str, np.ndarray], spec: OracleSpec, clips: np.ndarray,
            dtype=torch.float64, training: bool = False,
            taps: Optional[dict] = None, channels_last: bool = False) -> Dict[str, np.ndarray]:
    """X3D.call (model.py:113-127) on NDHWC clips.  Returns logits (fc2 output, :121),
    per-clip softmax, and the view-averaged probabilities the reference returns (:123-127).
    `taps`, when a dict, receives NCDHW intermediates keyed by layer for per-kernel parity.
    `channels_last`: keep the activations in torch's channels_last_3d memory format (the NDHWC
    layout the reference computes in) -- same arithmetic, the timing protocol of BASELINE.md
    section 3 for the CPU baseline."""
    x = torch.from_numpy(np.ascontiguousarray(clips)).to(dtype).permute(0, 4, 1, 2, 3)
    x = x.contiguous(memory_format=torch.channels_last_3d) if channels_last else x.contiguous()
    with torch.inference_mode():
        out = stem(W, x, spec, dtype)                                     # :114
        if taps is not None:
            taps["conv1"] = out
        for blk in spec.blocks:                                           # :115-116
            out = res_block(W, out, blk, spec, dtype, taps)
            if taps is not None:
                taps[f"stages/{blk[0]}/stage/layer_with_weights-{blk[1]}"] = out
        out = conv3d_valid(out, _k(W["conv5/layer_with_weights-0/kernel"], dtype), (1, 1, 1))
        out = F.relu(batchnorm_inference(out, W, "conv5/layer_with_weights-1", spec.bn_eps, dtype))
        out = global_avg_pool(out)                                        # :118
        if taps is not None:
            taps["pool5"] = out
        out = F.relu(conv3d_valid(out, _k(W["fc1/kernel"], dtype), (1, 1, 1)))    # :119
        # dropout (:120) is the identity at inference
        out = out.permute(0, 2, 3, 4, 1)                                  # Dense acts on last axis
        logits = out @ _v(W["fc2/kernel"], dtype) + _v(W["fc2/bias"], dtype)     # :121
        probs = torch.softmax(logits.to(torch.float32) if dtype != torch.float64 else logits,
                              dim=-1)                                     # :122 (fp32 softmax)
        nc = spec.num_classes
        if not training:                                                  # :123-126
            assert probs.shape[0] % spec.num_preds == 0, "batch must be a multiple of num_preds"
            avg = probs.reshape(-1, spec.num_preds, 1, 1, 1, nc).mean(1)
        else:
            avg = probs
        return {"logits": logits.reshape(-1, nc).numpy(),
                "clip_probs": probs.reshape(-1, nc).numpy(),
                "probs": avg.reshape(-1, nc).numpy()}                    # :127


def to_ndhwc(t: torch.Tensor) -> np.ndarray:
    return t.permute(0, 2, 3, 4, 1).contiguous().numpy()
