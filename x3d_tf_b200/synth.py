"""Seeded synthetic weights and clips.

The shipped checkpoints' data shards are not available (`.MISSING_LARGE_BLOBS:5-7` of the
reference) and there is no dataset, so benchmarks and parity tests run on synthetic weights laid
out under the exact checkpoint names/shapes and on synthetic clips
normalised like `utils.normalize` does (`utils.py:42-72`): (u8/255 - MEAN)/STD.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

from .arch import ArchSpec, variable_shapes


def synthetic_weights(arch: ArchSpec, seed: int = 1111, head_spread: float = 0.0) -> Dict[str, np.ndarray]:
    """Glorot-uniform kernels (Keras default initialiser), BN gamma~U[0.5,1.5],
    beta~N(0,0.1), moving_mean~N(0,0.1), moving_variance~U[0.5,1.5], small biases.

    `head_spread` > 0 (parity fixtures): the columns of `fc2/kernel` and the entries of `fc2/bias`
    are scaled by exp(N(0, head_spread)) per class.  With 400 near-Gaussian logits the top-1 /
    top-2 gap of a random-weight network is a few percent of the largest logit; a log-normal
    spread of the class scales makes the ranking heavy-tailed, so "identical top-1" becomes a
    decided question at the bf16 tolerance on every fixture clip."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shp in variable_shapes(arch).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            if len(shp) == 5:
                rf = shp[0] * shp[1] * shp[2]
                fan_in, fan_out = rf * shp[3], rf * shp[4]
                if shp[3] == 1 and rf > 1:         # channelwise: Keras counts groups' fan_out per group
                    fan_out = rf
            else:
                fan_in, fan_out = shp
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            # depthwise / stem kernels a little larger so signals do not vanish through 26 blocks
            a = rng.uniform(-lim, lim, size=shp)
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, size=shp)
        elif leaf == "moving_variance":
            a = rng.uniform(0.5, 1.5, size=shp)
        elif leaf in ("beta", "moving_mean"):
            a = rng.normal(0.0, 0.1, size=shp)
        elif leaf == "bias":
            a = rng.normal(0.0, 0.1, size=shp)
        else:
            raise KeyError(name)
        out[name] = a.astype(np.float32)
    if head_spread > 0.0:
        g = np.exp(np.random.default_rng(seed + 7919).normal(0.0, head_spread, size=out["fc2/bias"].shape))
        out["fc2/kernel"] = (out["fc2/kernel"] * g[None, :]).astype(np.float32)
        out["fc2/bias"] = (out["fc2/bias"] * g).astype(np.float32)
    return out


def synthetic_clips_u8(n: int, t: int, h: int, w: int, seed: int = 1111, c: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, t, h, w, c), dtype=np.uint8)


def normalize_clips(u8: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """`utils.normalize` (`utils.py:42-72`): x/255, -mean, /std; float32 result."""
    x = u8.astype(np.float32) / np.float32(255.0)
    return ((x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)).astype(np.float32)


def synthetic_clips(n: int, t: int, h: int, w: int, mean, std, seed: int = 1111) -> np.ndarray:
    return normalize_clips(synthetic_clips_u8(n, t, h, w, seed), mean, std)
