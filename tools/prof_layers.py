"""Runs individual layers at X3D-M shapes (for ncu captures and quick timing).
usage: python tools/prof_layers.py [dw|pw|stem|all] [--size 224] [--clips 8]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from x3d_tf_b200 import ops
from x3d_tf_b200.arch import same_pad

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="all")
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--T", type=int, default=16)
ap.add_argument("--se", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
N, T, S0 = a.clips, a.T, a.size // 2
g = torch.Generator(device=dev); g.manual_seed(0)
def rnd(*shape, dtype=torch.bfloat16): return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32).to(dtype)
def timeit(name, fn, bytes_):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f"{name:44s} {ms:8.3f} ms  {bytes_ / ms / 1e6:8.1f} GB/s", flush=True)

stages = [(S0, 24, 54, 24), (S0 // 2, 24, 108, 48), (S0 // 4, 48, 216, 96), (S0 // 8, 96, 432, 192)]
pad8 = lambda c: (c + 7) // 8 * 8
for si, (Hin, cin, inner, cout) in enumerate(stages):
    ci = pad8(inner)
    for stride in (2, 1):
        H = Hin if stride == 2 else Hin // 2
        Ho = H // stride
        if a.what in ("dw", "all"):
            x = rnd(N, T, H, H, ci); w = rnd(27, ci, dtype=torch.float32); b = rnd(ci, dtype=torch.float32)
            _, ph, _ = same_pad(H, 3, stride)
            for se in (bool(a.se),):
                timeit(f"dw s{si+2} {H}x{H}x{ci} stride{stride} se{int(se)}", lambda: ops.dw_fwd(x, w, b, stride, ph, ph, se),
                       (x.numel() + N * T * Ho * Ho * ci) * 2)
        if a.what in ("dwp", "all") and ops.dw_planar_supported(T, H, H, ci, stride) > 0:
            x = rnd(N, T, H, H, ci); w = rnd(27, ci, dtype=torch.float32); b = rnd(ci, dtype=torch.float32)
            _, ph, _ = same_pad(H, 3, stride)
            taps = ops.dw_planar_taps(w, b)
            timeit(f"dwp s{si+2} {H}x{H}x{ci} stride{stride} se{a.se}", lambda: ops.dw_planar_fwd(x, taps, stride, ph, ph, bool(a.se)),
                   (x.numel() + N * T * Ho * Ho * ci) * 2)
            timeit(f"  (dw_tma)", lambda: ops.dw_fwd(x, w, b, stride, ph, ph, bool(a.se)), (x.numel() + N * T * Ho * Ho * ci) * 2)
        if a.what in ("ab", "ab2", "all"):
            cin_s = pad8(cin if stride == 2 else cout)
            x = rnd(N, T, H, H, cin_s); w = rnd(27, ci, dtype=torch.float32); b = rnd(ci, dtype=torch.float32)
            wa = rnd((ci + 15) // 16 * 16, (cin_s + 63) // 64 * 64); ba = rnd(ci, dtype=torch.float32)
            _, ph, _ = same_pad(H, 3, stride)
            if a.what != "ab2" and ops.expand_dw_supported(T, H, H, cin_s, ci, stride) > 0:
                timeit(f"ab s{si+2} {H}x{H} {cin_s}->{ci} stride{stride}", lambda: ops.expand_dw_fwd(x, wa, ba, w, b, stride, ph, ph, True),
                       (x.numel() + N * T * Ho * Ho * ci) * 2)
            if a.what in ("ab", "ab2", "all") and ops.expand_dw2_supported(T, H, H, cin_s, ci, stride) > 0:
                timeit(f"ab2 s{si+2} {H}x{H} {cin_s}->{ci} stride{stride}", lambda: ops.expand_dw2_fwd(x, wa, ba, w, b, stride, ph, ph, bool(a.se)),
                       (x.numel() + N * T * Ho * Ho * ci) * 2)
            M = N * T * H * H
            timeit(f"  (a alone: M={M} K={cin_s} N={ci})", lambda: ops.pw_tc_fwd(x.view(M, cin_s), wa, ba, M=M, K=cin_s, Nc=ci, relu=True),
                   (x.numel() + M * ci) * 2)
            xx = rnd(N, T, H, H, ci)
            timeit(f"  (b alone)", lambda: ops.dw_fwd(xx, w, b, stride, ph, ph, True), (xx.numel() + N * T * Ho * Ho * ci) * 2)
        if a.what in ("pw", "all") and stride == 1:
            M = N * T * H * H
            xa = rnd(M, pad8(cin if False else cout)); K = xa.shape[1]
            wp = rnd((ci + 15) // 16 * 16, (K + 63) // 64 * 64); bias = rnd(ci, dtype=torch.float32)
            timeit(f"pw-a tc s{si+2} M={M} K={K} N={ci}", lambda: ops.pw_tc_fwd(xa, wp, bias, M=M, K=K, Nc=ci, relu=True),
                   (xa.numel() + M * ci) * 2)
            xb = rnd(M, ci); wp2 = rnd((K + 15) // 16 * 16, (ci + 63) // 64 * 64); bias2 = rnd(K, dtype=torch.float32)
            res = rnd(M, K); sev = torch.rand(N, ci, device=dev)
            timeit(f"pw-c tc s{si+2} M={M} K={ci} N={K} (se+swish+res)",
                   lambda: ops.pw_tc_fwd(xb, wp2, bias2, M=M, K=ci, Nc=K, residual=res, se=sev, rows_per_clip=T * H * H, swish=True, relu=True),
                   (xb.numel() + 2 * M * K) * 2)
            timeit(f"pw-c tc s{si+2} M={M} K={ci} N={K} (plain+res)",
                   lambda: ops.pw_tc_fwd(xb, wp2, bias2, M=M, K=ci, Nc=K, residual=res, relu=True),
                   (xb.numel() + 2 * M * K) * 2)
if a.what in ("stem", "all"):
    x = rnd(N, T, a.size, a.size, 3); ws = rnd(27, 24, dtype=torch.float32); wt = rnd(5, 24, dtype=torch.float32); b = rnd(24, dtype=torch.float32)
    timeit(f"stem simt {a.size}", lambda: ops.stem_fwd(x, ws, wt, b, torch.bfloat16), (x.numel() + N * T * S0 * S0 * 24) * 2)
    wc = rnd(5, 4, 32, 8)
    timeit(f"stem tcgen05 {a.size}", lambda: ops.stem_tc_fwd(x, wc, b), (x.numel() + N * T * S0 * S0 * 24) * 2)
