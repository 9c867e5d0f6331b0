"""fp32 pointwise GEMM (x3d_pw_fwd, 3xTF32 tensor-core path) at the training step's layer shapes:
CUDA-event time and algorithmic GB/s per shape.  usage: python tools/prof_pw32.py [--reps 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from x3d_tf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", type=str, default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
B = 32 * 16
shapes = [("s2 expand", B * 56 * 56, 24, 56), ("s2 project", B * 56 * 56, 56, 24),
          ("s3 expand", B * 28 * 28, 48, 112), ("s3 project", B * 28 * 28, 112, 48),
          ("s4 expand", B * 14 * 14, 96, 216), ("s4 project", B * 14 * 14, 216, 96),
          ("s5 expand", B * 7 * 7, 192, 432), ("s5 project", B * 7 * 7, 432, 192)]
for name, M, K, N in shapes:
    if a.only and a.only not in name:
        continue
    x = torch.randn(M, K, generator=g, device=dev)
    w = torch.randn(K, N, generator=g, device=dev) * 0.1
    b = torch.randn(N, generator=g, device=dev)
    for _ in range(2):
        y = ops.pw_fwd(x, w, b, M=M, K=K, Nc=N, out_dtype=torch.float32, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        y = ops.pw_fwd(x, w, b, M=M, K=K, Nc=N, out_dtype=torch.float32, relu=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    ref = torch.relu(x[:4096].double() @ w.double() + b.double())
    err = float((y[:4096].double() - ref).abs().max() / ref.abs().max())
    print(f"{name:12s} M={M} K={K} N={N}: {ms * 1e3:8.1f} us  {M * (K + N) * 4 / ms / 1e6:7.0f} GB/s  "
          f"{2 * M * K * N / ms / 1e9:6.1f} TFLOP/s  rel err {err:.1e}")
