"""Second, independent CPU statement (numpy, float64, explicit loops over taps) of the layers
whose TensorFlow semantics differ from torch defaults.  TEST INFRASTRUCTURE ONLY (see
`oracle/x3d_oracle.py` header).  Used to cross-check the torch oracle; operates on NDHWC arrays
with TF-layout kernels, written directly from the TF definitions, sharing no code with it.

  * channelwise 3x3x3 conv, stride (1,s,s), padding='same'   -- model.py:259-267
  * stem: pad(0,1,1) -> 1x3x3 s(1,2,2) valid -> pad(2,0,0) -> 5x1x1 channelwise valid  -- model.py:202-206
  * 1x1x1 strided 'valid' conv (shortcut)                    -- model.py:360-367
"""
from __future__ import annotations

import numpy as np


def _same(in_size, k, s):
    out = (in_size + s - 1) // s
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2


def channelwise_conv_same(x: np.ndarray, kernel: np.ndarray, stride: int) -> np.ndarray:
    """x [N,T,H,W,C]; kernel [3,3,3,1,C] (DHWIO, groups=C)."""
    x = x.astype(np.float64)
    N, T, H, W, C = x.shape
    kd, kh, kw = kernel.shape[:3]
    To, pt = _same(T, kd, 1)
    Ho, ph = _same(H, kh, stride)
    Wo, pw = _same(W, kw, stride)
    out = np.zeros((N, To, Ho, Wo, C), np.float64)
    for dt in range(kd):
        for dh in range(kh):
            for dw in range(kw):
                wv = kernel[dt, dh, dw, 0, :].astype(np.float64)
                for to in range(To):
                    ti = to - pt + dt
                    if ti < 0 or ti >= T:
                        continue
                    # output rows whose input row is in range
                    hs = [ho for ho in range(Ho) if 0 <= ho * stride - ph + dh < H]
                    ws = [wo for wo in range(Wo) if 0 <= wo * stride - pw + dw < W]
                    if not hs or not ws:
                        continue
                    hi = np.array(hs) * stride - ph + dh
                    wi = np.array(ws) * stride - pw + dw
                    out[:, to, hs[0]:hs[-1] + 1, ws[0]:ws[-1] + 1, :] += \
                        x[:, ti][:, hi][:, :, wi] * wv
    return out


def stem_convs(x: np.ndarray, ks: np.ndarray, kt: np.ndarray) -> np.ndarray:
    """x [N,T,H,W,3]; ks [1,3,3,3,C]; kt [kT,1,1,1,C].  Returns conv_t(pad(conv_s(pad(x))))."""
    x = x.astype(np.float64)
    N, T, H, W, Ci = x.shape
    C = ks.shape[-1]
    xp = np.zeros((N, T, H + 2, W + 2, Ci), np.float64)
    xp[:, :, 1:H + 1, 1:W + 1] = x
    Ho = (H + 2 - 3) // 2 + 1
    Wo = (W + 2 - 3) // 2 + 1
    s = np.zeros((N, T, Ho, Wo, C), np.float64)
    for dh in range(3):
        for dw in range(3):
            patch = xp[:, :, dh:dh + 2 * (Ho - 1) + 1:2, dw:dw + 2 * (Wo - 1) + 1:2, :]
            s += np.einsum("nthwi,ic->nthwc", patch, ks[0, dh, dw].astype(np.float64))
    kT = kt.shape[0]
    pt = kT // 2
    sp = np.zeros((N, T + 2 * pt, Ho, Wo, C), np.float64)
    sp[:, pt:pt + T] = s
    out = np.zeros_like(s)
    for dt in range(kT):
        out += sp[:, dt:dt + T] * kt[dt, 0, 0, 0, :].astype(np.float64)
    return out


def pointwise_conv_valid(x: np.ndarray, kernel: np.ndarray, stride: int = 1) -> np.ndarray:
    """x [N,T,H,W,Ci]; kernel [1,1,1,Ci,Co]; stride on H and W, 'valid'."""
    xs = x.astype(np.float64)[:, :, ::stride, ::stride, :]
    return xs @ kernel[0, 0, 0].astype(np.float64)


def tf32x3_matmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Arithmetic model of the 3xTF32 pointwise GEMM (x3d_tf_b200/csrc/x3d_simt.cu: split_tf32 +
    mma.sync.m16n8k8.tf32; used for the fp32 pointwise convs, reference model.py:246-258,292-303):
    every fp32 operand is split into hi = x rounded to TF32 (10 mantissa bits, half away from zero,
    integer add + mask) and lo = x - hi (exact in fp32) pre-biased by half a TF32 ulp and truncated to
    TF32 as the tensor core does with the low 13 bits; the product is lo.hi + hi.lo + hi.hi.  The sums
    are taken in float64 here, so the result isolates the error of the SPLIT (the dropped lo.lo term
    and the rounding of lo); the kernel's fp32 accumulation adds ordinary fp32 rounding on top."""
    def split(x):
        x = np.ascontiguousarray(x, np.float32)
        bits = x.view(np.uint32)
        hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
        lo = (x - hi).astype(np.float32)
        lo_t = ((lo.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
        return hi.astype(np.float64), lo_t.astype(np.float64)
    ah, al = split(a)
    bh, bl = split(b)
    return al @ bh + ah @ bl + ah @ bh
