"""One training step under the CUDA profiler range (for `ncu --profile-from-start off`).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
           --log-file gpurun_out/train_launches.csv python tools/prof_train.py [--clips 32]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import device_clips
from x3d_tf_b200.arch import build_arch
from x3d_tf_b200.config import get_config
from x3d_tf_b200.synth import synthetic_weights
from x3d_tf_b200.training import X3DTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=32)
ap.add_argument("--size", type=int, default=224)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = get_config("X3D_M")
tr = X3DTrainer(cfg, device=dev).load(synthetic_weights(build_arch(cfg), seed=1111))
x = device_clips(a.clips, 16, a.size, cfg, torch.float32, dev, seed=1)
labels = torch.randint(0, 400, (a.clips,), device=dev, dtype=torch.int32)
tr.step(x, labels, 0.01)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(x, labels, 0.01)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
