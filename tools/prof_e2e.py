"""Where does predict() lose time against back-to-back graph replays?  (diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from x3d_tf_b200 import ops

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
model, cfg, arch, _ = bench.build_model("m256x10", graph=True)
clips, T, S = 80, 16, 256
u8 = torch.randint(0, 256, (clips, T, S, S, 3), dtype=torch.uint8).pin_memory()
xd = u8.to(dev)
model(xd); model(xd); torch.cuda.synchronize()

def timed(fn, n=8):
    fn(2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

def replay_only(n):
    for _ in range(n): model(xd)
def predict_host(n):
    for _ in model.predict(u8 for _ in range(n)): pass
def h2d_only(n):
    for _ in range(n): xd.copy_(u8, non_blocking=True)
def predict_device_src(n):          # same loop, source already on the device (no PCIe)
    for _ in model.predict(xd for _ in range(n)): pass

for name, fn in [("graph replay, device input", replay_only), ("H2D 252 MB only", h2d_only),
                 ("predict(host uint8)", predict_host), ("predict(device uint8 source)", predict_device_src)]:
    gpu, wall = timed(fn)
    print(f"{name:34s} {gpu:8.3f} ms/step (events)  {wall:8.3f} ms/step (wall)", flush=True)
