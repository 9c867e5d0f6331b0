"""CPU ORACLE for one X3D training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, with torch-CPU autograd in float64, what Keras `fit` does per replica for the reference
(train.py:85-152 -> model.py with training=True):
  forward with BatchNormalization in batch-statistics mode (biased variance, eps 1e-5), dropout
  (mask supplied by the caller so the check is deterministic), softmax, SparseCategoricalCrossentropy
  on the probabilities (clipped to [1e-7, 1-1e-7]) averaged over the batch, plus the L2 regulariser
  WEIGHT_DECAY * sum(w^2) of every Conv3D/Dense kernel except se_fc1 (model.py:47, 278-283);
  gradients by autograd; SGD(nesterov=True, momentum) update (train.py:88-92):
      v <- mu*v - lr*g ;  w <- w + mu*v - lr*g
  moving statistics: moving <- 0.9*moving + 0.1*batch (configs/default.py:43).

PARITY UNPINNED against TensorFlow itself (TF cannot run here); pinned against the forward oracle
(`x3d_oracle.forward(training=False)` agrees when the moving statistics equal the batch statistics)
and by torch autograd being an independent derivation of every backward kernel.
Only `tests/` may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import x3d_oracle as O

REGULARISED_SUFFIX = "/kernel"


def is_regularised(name: str) -> bool:
    """kernel_regularizer is set on every Conv3D/Dense except se_fc1 (model.py:278-283)."""
    return name.endswith(REGULARISED_SUFFIX) and "/se_fc1/" not in name


def _bn_train(x, P, prefix, eps, stats):
    g, b = P[prefix + "/gamma"], P[prefix + "/beta"]
    mean = x.mean(dim=(0, 2, 3, 4))
    var = x.var(dim=(0, 2, 3, 4), unbiased=False)
    stats[prefix] = (mean.detach(), var.detach())
    sh = (1, -1, 1, 1, 1)
    return (x - mean.view(sh)) * torch.rsqrt(var.view(sh) + eps) * g.view(sh) + b.view(sh)


def _k(t):           # DHWIO -> OIDHW
    return t.permute(4, 3, 0, 1, 2)


def forward_train(P: Dict[str, torch.Tensor], spec: O.OracleSpec, x: torch.Tensor,
                  dropout_mask: Optional[torch.Tensor], stats: dict, taps: Optional[dict] = None,
                  relu_masks: Optional[list] = None):
    """x: NCDHW.  Returns logits [N, classes].  `taps` (a dict) receives intermediate activations
    with retain_grad() set, for localising a backward mismatch."""
    def tap(name, t):
        if taps is not None:
            t.retain_grad()
            taps[name] = t
        return t
    masks = None if relu_masks is None else list(relu_masks)

    def relu(t):
        """F.relu, or -- when the caller supplies the sign pattern the implementation under test
        actually used (one [rows, channels] bool array per ReLU, in execution order) -- t * mask.
        The two differ only where |t| is below the fp32 forward error, but there the float64 and
        fp32 runs pick different branches of the kink and their gradients are not comparable."""
        if masks is None:
            return F.relu(t)
        m = torch.from_numpy(np.asarray(masks.pop(0))).to(t.dtype)
        c = t.shape[1]
        m = m.reshape(t.shape[0], *t.shape[2:], -1)[..., :c].permute(0, 4, 1, 2, 3) if t.dim() == 5 \
            else m.reshape(t.shape[0], -1)[:, :c]
        return t * m
    eps = spec.bn_eps
    out = F.pad(x, (1, 1, 1, 1, 0, 0))
    out = F.conv3d(out, _k(P["conv1/conv_s/kernel"]), None, stride=(1, 2, 2))
    tp = spec.temp_filter // 2
    out = F.pad(out, (0, 0, 0, 0, tp, tp))
    out = F.conv3d(out, _k(P["conv1/conv_t/kernel"]), None, groups=spec.conv1_dim)
    out = relu(_bn_train(out, P, "conv1/bn", eps, stats))
    for (s, j, cin, inner, cout, stride, se, shortcut) in spec.blocks:
        p = f"stages/{s}/stage/layer_with_weights-{j}"
        q = p + "/bottleneck"
        xin = out
        out = O.conv3d_same(xin, _k(P[q + "/a/kernel"]), (1, 1, 1), 1)
        out = relu(_bn_train(out, P, q + "/bn_a", eps, stats))
        out = O.conv3d_same(out, _k(P[q + "/b/kernel"]), (1, stride, stride), inner)
        out = _bn_train(out, P, q + "/bn_b", eps, stats)
        if se:
            m = out.mean(dim=(2, 3, 4), keepdim=True)
            z = F.relu(F.conv3d(m, _k(P[q + "/se_fc1/kernel"]), P[q + "/se_fc1/bias"]))
            sc = torch.sigmoid(F.conv3d(z, _k(P[q + "/se_fc2/kernel"]), P[q + "/se_fc2/bias"]))
            out = out * sc
        out = out * torch.sigmoid(out)
        out = O.conv3d_same(out, _k(P[q + "/c/kernel"]), (1, 1, 1), 1)
        out = _bn_train(out, P, q + "/bn_c", eps, stats)
        if shortcut:
            res = F.conv3d(xin, _k(P[p + "/residual/kernel"]), None, stride=(1, stride, stride))
            res = _bn_train(res, P, p + "/bn_r", eps, stats)
        else:
            res = xin
        out = tap(p, relu(res + out))
    out = tap("conv5_pre", F.conv3d(out, _k(P["conv5/layer_with_weights-0/kernel"]), None))
    out = tap("conv5", relu(_bn_train(out, P, "conv5/layer_with_weights-1", eps, stats)))
    out = tap("pool5", out.mean(dim=(2, 3, 4), keepdim=True))
    out = F.conv3d(out, _k(P["fc1/kernel"]), None).reshape(out.shape[0], -1)
    out = tap("fc1", relu(out))
    if dropout_mask is not None:
        out = out * dropout_mask
    return out @ P["fc2/kernel"] + P["fc2/bias"]


def train_step(W: Dict[str, np.ndarray], spec: O.OracleSpec, clips: np.ndarray, labels: np.ndarray,
               *, lr: float, momentum: float = 0.9, weight_decay: float = 5e-5,
               bn_momentum: float = 0.9, velocity: Optional[Dict[str, np.ndarray]] = None,
               dropout_mask: Optional[np.ndarray] = None, world: int = 1, dtype=torch.float64,
               taps: Optional[dict] = None, relu_masks: Optional[list] = None,
               optimizer: str = "sgd", adam_state: Optional[dict] = None):
    """One replica's step.  `world` > 1 scales the data loss by 1/world as Keras does under
    MirroredStrategy (gradients are then summed across replicas by the caller).
    `optimizer="adam"` applies Keras' Adam instead (train.py:93-95; tensorflow==2.4.1
    `optimizer_v2/adam.py`: beta_1 0.9, beta_2 0.999, epsilon 1e-7, lr_t = lr sqrt(1-b2^t)/(1-b1^t));
    `adam_state` = dict(t=steps done, m={...}, v={...}).
    Returns dict(loss, grads, weights, velocity, logits[, adam_v])."""
    P = {}
    for k, v in W.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        trainable = not (k.endswith("moving_mean") or k.endswith("moving_variance"))
        P[k] = t.requires_grad_(trainable)
    x = torch.from_numpy(np.ascontiguousarray(clips)).to(dtype).permute(0, 4, 1, 2, 3).contiguous()
    mask = None if dropout_mask is None else torch.from_numpy(dropout_mask).to(dtype)
    stats: dict = {}
    logits = forward_train(P, spec, x, mask, stats, taps, relu_masks)
    probs = torch.softmax(logits, dim=-1)
    y = torch.from_numpy(np.asarray(labels, np.int64))
    py = probs.gather(1, y[:, None])[:, 0].clamp(1e-7, 1 - 1e-7)
    data_loss = (-torch.log(py)).mean() / world
    reg = sum((weight_decay * (P[k] ** 2).sum()) for k in P if is_regularised(k)) / world
    loss = data_loss + reg
    names = [k for k in P if P[k].requires_grad]
    grads = torch.autograd.grad(loss, [P[k] for k in names])   # retain_grad() hooks of `taps` fire here too
    G = {k: g.detach().numpy() for k, g in zip(names, grads)}
    newW, newV, newV2 = {}, {}, {}
    for k in W:
        if k in G and optimizer == "adam":
            b1, b2, eps = 0.9, 0.999, 1e-7
            st = adam_state or {"t": 0, "m": {}, "v": {}}
            t = st["t"] + 1
            m0 = st["m"].get(k, np.zeros_like(G[k])).astype(np.float64)
            s0 = st["v"].get(k, np.zeros_like(G[k])).astype(np.float64)
            m1 = b1 * m0 + (1 - b1) * G[k]
            s1 = b2 * s0 + (1 - b2) * G[k] ** 2
            lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
            newV[k], newV2[k] = m1, s1
            newW[k] = W[k].astype(np.float64) - lr_t * m1 / (np.sqrt(s1) + eps)
        elif k in G:
            v0 = np.zeros_like(G[k]) if velocity is None else velocity[k].astype(np.float64)
            v1 = momentum * v0 - lr * G[k]
            newV[k] = v1
            newW[k] = W[k].astype(np.float64) + momentum * v1 - lr * G[k]
        else:
            newW[k] = W[k].astype(np.float64)
    for prefix, (mean, var) in stats.items():
        newW[prefix + "/moving_mean"] = bn_momentum * W[prefix + "/moving_mean"].astype(np.float64) + \
            (1 - bn_momentum) * mean.numpy()
        newW[prefix + "/moving_variance"] = bn_momentum * W[prefix + "/moving_variance"].astype(np.float64) + \
            (1 - bn_momentum) * var.numpy()
    return {"loss": float(data_loss.detach() * world), "grads": G, "weights": newW, "velocity": newV,
            "adam_v": newV2, "logits": logits.detach().numpy()}
