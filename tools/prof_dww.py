"""Channelwise backward-filter (x3d_dw_wgrad) at the training step's layer shapes: CUDA-event time.
usage: python tools/prof_dww.py [--only s5] [--reps 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x3d_tf_b200._lib import lib

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", type=str, default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for name, H, C in [("s2", 56, 56), ("s3", 28, 112), ("s4", 14, 216), ("s5", 7, 432)]:
    if a.only and a.only != name:
        continue
    N, T = 32, 16
    x = torch.randn(N, T, H, H, C, device=dev); dy = torch.randn(N, T, H, H, C, device=dev)
    dw = torch.zeros(27, C, dtype=torch.float64, device=dev)
    def run():
        assert lib().x3d_dw_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, T, H, H, C, 1, 1, 1, st) == 0
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f"{name}: [{N},{T},{H},{H},{C}] {ms * 1e3:8.1f} us  {2 * x.numel() * 4 / ms / 1e6:7.0f} GB/s (x + dy once)")
