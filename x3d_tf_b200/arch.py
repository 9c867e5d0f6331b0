"""Host-side derivation of the X3D graph from a config.

Restates what `X3D.__init__` computes before it builds any layer (reference `model.py:24-76`)
and the width/depth rounding of `utils.py:7-40`, plus the shape arithmetic TensorFlow applies
to the layers on the path (SAME padding of the strided channelwise conv, `valid` strided
shortcut, explicit symmetric stem padding; `model.py:161-175,259-267,360-367`).

Everything here is plain integer arithmetic on the host; the CUDA kernels receive the results
(channel counts, strides, pad-before values) as launch parameters.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Tuple


def round_width(width, multiplier, min_depth=8, divisor=8):
    """Channel rounding rule (reference `utils.py:7-30`)."""
    if not multiplier:
        return width
    width *= multiplier
    min_depth = min_depth or divisor
    new_filters = max(min_depth, int(width + divisor / 2) // divisor * divisor)
    if new_filters < 0.9 * width:
        new_filters += divisor
    return int(new_filters)


def round_repeats(repeats, multiplier):
    """Depth rounding rule (reference `utils.py:32-40`)."""
    if not multiplier:
        return repeats
    return int(math.ceil(multiplier * repeats))


def same_pad(in_size: int, kernel: int, stride: int) -> Tuple[int, int, int]:
    """TensorFlow `padding='same'`: returns (out_size, pad_before, pad_after)."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + kernel - in_size, 0)
    before = total // 2
    return out, before, total - before


@dataclass
class BlockSpec:
    stage: int            # 0..3  (res_stage_2..5)
    index: int            # index inside the stage (checkpoint `layer_with_weights-{index}`)
    block_index: int      # value the reference passes as Bottleneck.block_index (global, 1-based)
    cin: int
    cinner: int
    cout: int
    stride: int
    se_width: int         # 0 = no squeeze-excitation
    has_shortcut: bool

    @property
    def has_se(self) -> bool:
        return self.se_width > 0

    @property
    def prefix(self) -> str:
        return f"stages/{self.stage}/stage/layer_with_weights-{self.index}"


@dataclass
class ArchSpec:
    stem_channels: int
    temp_filter: int
    blocks: List[BlockSpec]
    stage_dims: List[Tuple[int, int, int, int]]   # (depth, in, inner, out) per stage
    conv5_channels: int
    fc1_channels: int
    num_classes: int
    bn_eps: float
    bn_momentum: float
    num_preds: int
    dropout_rate: float
    weight_decay: float
    in_channels: int = 3
    se_ratio: float = 0.0625

    def stage_blocks(self, s: int) -> List[BlockSpec]:
        return [b for b in self.blocks if b.stage == s]


def se_enabled(block_index: int) -> bool:
    """`model.py:275,311`: SE iff (block_index + 1) is even."""
    return (block_index + 1) % 2 == 0


def build_arch(cfg, first_block_index: int = 1, se_ratio: float = 0.0625) -> ArchSpec:
    """Derive the whole graph.  `first_block_index` is the value `ResBlock._block_index` will
    have after its increment for the first block (1 in a fresh process; `model.py:326,351,378`)."""
    net = cfg.NETWORK
    if net.SCALE_RES2:
        conv1_dim = round_width(net.C1_CHANNELS, net.WIDTH_FACTOR)
        mult = 1
    else:
        conv1_dim = round_width(net.C1_CHANNELS, 2)
        mult = 2
    base = net.C1_CHANNELS * mult
    basis = [[1, base], [2, round_width(base, 2)], [5, round_width(base, 4)],
             [3, round_width(base, 8)]]
    blocks: List[BlockSpec] = []
    stage_dims = []
    out_dim = conv1_dim
    g = first_block_index
    for s, (bd, bc) in enumerate(basis):
        in_dim = out_dim
        out_dim = round_width(bc, net.WIDTH_FACTOR)
        inner = int(out_dim * net.BOTTLENECK_WIDTH_FACTOR)
        depth = round_repeats(bd, net.DEPTH_FACTOR)
        stage_dims.append((depth, in_dim, inner, out_dim))
        for i in range(depth):
            cin = in_dim if i == 0 else out_dim
            stride = 2 if i == 0 else 1
            blocks.append(BlockSpec(
                stage=s, index=i, block_index=g, cin=cin, cinner=inner, cout=out_dim,
                stride=stride,
                se_width=round_width(inner, se_ratio) if se_enabled(g) else 0,
                has_shortcut=(cin != out_dim or stride != 1)))
            g += 1
    return ArchSpec(
        stem_channels=conv1_dim, temp_filter=net.C1_TEMP_FILTER, blocks=blocks,
        stage_dims=stage_dims, conv5_channels=stage_dims[-1][2], fc1_channels=2048,
        num_classes=net.NUM_CLASSES, bn_eps=float(net.BN.EPS),
        bn_momentum=float(net.BN.MOMENTUM),
        num_preds=cfg.TEST.NUM_TEMPORAL_VIEWS * cfg.TEST.NUM_SPATIAL_CROPS,
        dropout_rate=float(net.DROPOUT_RATE), weight_decay=float(net.WEIGHT_DECAY),
        in_channels=cfg.DATA.NUM_INPUT_CHANNELS, se_ratio=se_ratio)


_BN_VARS = ("gamma", "beta", "moving_mean", "moving_variance")


def variable_shapes(arch: ArchSpec) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every model variable under its TF-checkpoint attribute path (without the
    `/.ATTRIBUTES/VARIABLE_VALUE` suffix) with its shape; conv kernels are DHWIO."""
    v: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def bn(prefix, c):
        for n in _BN_VARS:
            v[f"{prefix}/{n}"] = (c,)

    c1 = arch.stem_channels
    v["conv1/conv_s/kernel"] = (1, 3, 3, arch.in_channels, c1)
    v["conv1/conv_t/kernel"] = (arch.temp_filter, 1, 1, 1, c1)
    bn("conv1/bn", c1)
    for b in arch.blocks:
        p = b.prefix
        if b.has_shortcut:
            v[f"{p}/residual/kernel"] = (1, 1, 1, b.cin, b.cout)
            bn(f"{p}/bn_r", b.cout)
        q = f"{p}/bottleneck"
        v[f"{q}/a/kernel"] = (1, 1, 1, b.cin, b.cinner)
        bn(f"{q}/bn_a", b.cinner)
        v[f"{q}/b/kernel"] = (3, 3, 3, 1, b.cinner)
        bn(f"{q}/bn_b", b.cinner)
        if b.has_se:
            v[f"{q}/se_fc1/kernel"] = (1, 1, 1, b.cinner, b.se_width)
            v[f"{q}/se_fc1/bias"] = (b.se_width,)
            v[f"{q}/se_fc2/kernel"] = (1, 1, 1, b.se_width, b.cinner)
            v[f"{q}/se_fc2/bias"] = (b.cinner,)
        v[f"{q}/c/kernel"] = (1, 1, 1, b.cinner, b.cout)
        bn(f"{q}/bn_c", b.cout)
    cl = arch.blocks[-1].cout
    v["conv5/layer_with_weights-0/kernel"] = (1, 1, 1, cl, arch.conv5_channels)
    bn("conv5/layer_with_weights-1", arch.conv5_channels)
    v["fc1/kernel"] = (1, 1, 1, arch.conv5_channels, arch.fc1_channels)
    v["fc2/kernel"] = (arch.fc1_channels, arch.num_classes)
    v["fc2/bias"] = (arch.num_classes,)
    return v


def is_trainable(name: str) -> bool:
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


def param_counts(arch: ArchSpec) -> Dict[str, int]:
    """Parameter totals in the grouping Keras `summary()` prints (`models/*/X3D_*.txt`)."""
    shapes = variable_shapes(arch)
    out = {"conv_1": 0, "conv_5": 0, "fc_1": 0, "fc_2": 0, "total": 0, "trainable": 0}
    for s in range(4):
        out[f"res_stage_{s + 2}"] = 0
    for name, shp in shapes.items():
        n = math.prod(shp)
        out["total"] += n
        if is_trainable(name):
            out["trainable"] += n
        if name.startswith("conv1/"):
            out["conv_1"] += n
        elif name.startswith("stages/"):
            out[f"res_stage_{int(name.split('/')[1]) + 2}"] += n
        elif name.startswith("conv5/"):
            out["conv_5"] += n
        elif name.startswith("fc1/"):
            out["fc_1"] += n
        elif name.startswith("fc2/"):
            out["fc_2"] += n
    out["non_trainable"] = out["total"] - out["trainable"]
    return out


@dataclass
class LevelShape:
    T: int
    H: int
    W: int

    @property
    def P(self) -> int:
        return self.T * self.H * self.W


@dataclass
class ShapePlan:
    """Spatial extents at every level for one input size, plus the SAME pad-before of the
    stride-2 channelwise conv that enters each stage."""
    input: LevelShape
    stem: LevelShape
    stages: List[LevelShape] = field(default_factory=list)
    pads: List[Tuple[int, int]] = field(default_factory=list)   # (pad_before_h, pad_before_w)


def plan_shapes(arch: ArchSpec, T: int, H: int, W: int) -> ShapePlan:
    # stem: explicit (1,1) pad + 3x3 valid stride 2  (model.py:161-166,178-184)
    sh, sw = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    plan = ShapePlan(LevelShape(T, H, W), LevelShape(T, sh, sw))
    h, w = sh, sw
    for _ in range(4):
        oh, ph, _ = same_pad(h, 3, 2)
        ow, pw, _ = same_pad(w, 3, 2)
        plan.stages.append(LevelShape(T, oh, ow))
        plan.pads.append((ph, pw))
        h, w = oh, ow
    return plan
