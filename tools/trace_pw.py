"""Epilogue timeline of the pointwise kernel (needs the library built with -DX3D_PW_TRACE:
   nvcc ... -DX3D_PW_TRACE -c x3d_pw_tc.cu, see tools/gpu_trace.sh).  Prints, for one epilogue group
   of one CTA, the clocks between the marks of consecutive (sub-)tiles."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from x3d_tf_b200 import ops, _lib

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
def rnd(*shape, dtype=torch.bfloat16): return torch.randn(*shape, generator=g, device=dev).to(dtype)
cases = {"a_s2": (40 * 16 * 64 * 64, 24, 56, False), "c_s2_plain": (40 * 16 * 64 * 64, 56, 24, True),
         "a_s3": (40 * 16 * 32 * 32, 48, 112, False)}
names = ["loop top", "t_full passed", "drain+math+STS done", "t_empty arrive + fence", "wait_read", "bar.sync", "TMA store issued"]
for name, (M, K, N, res) in cases.items():
    x = rnd(M, K); w = rnd((N + 15) // 16 * 16, (K + 63) // 64 * 64); b = rnd(N, dtype=torch.float32)
    r = rnd(M, N) if res else None
    for _ in range(2):
        ops.pw_tc_fwd(x, w, b, M=M, K=K, Nc=N, residual=r, relu=True)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (8 * 256))()
    assert _lib.lib().x3d_pw_trace_dump(buf) == 0
    full = np.frombuffer(buf, dtype=np.int64).reshape(256, 8)
    full = full[(full[:, 0] > 0)][8:120]
    print(f"   (t_full passed -> first tcgen05.ld returned: {(full[:, 7] - full[:, 1]).mean():.0f} clk)")
    t = full[:, :7]
    t = t[(t[:, 0] > 0)]
    t = np.vstack([np.zeros((8, 7), np.int64), t])
    t = t[8:120]                                   # steady state
    d = np.diff(t, axis=1)
    period = np.diff(t[:, 0])
    print(f"== {name}: M={M} K={K} N={N}: tile period of this group {period.mean():.0f} clk (min {period.min()}, max {period.max()})")
    for i in range(6):
        print(f"   {names[i]:26s} -> {names[i + 1]:26s} {d[:, i].mean():8.0f} clk")
