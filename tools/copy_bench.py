"""Pinned host -> device copy bandwidth with every rank copying at the same time (the ceiling of the
end-to-end clips/s figure at N GPUs: each step moves one batch of clips over PCIe).
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/copy_bench.py [--mb 252]"""
import argparse, json, os
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=251.7, help="bytes per copy in MB (80 uint8 clips of 16x256x256x3 = 251.7)")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(a.mb * 1e6)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
devbuf = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    devbuf.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    devbuf.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
allms = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allms, t)
else:
    allms = [t]
if rank == 0:
    per = [n / (float(x.item()) * 1e-3) / 1e9 for x in allms]
    print(json.dumps({"n_gpus": world, "bytes_per_copy": n, "ms_per_copy_max": max(float(x.item()) for x in allms),
                      "GBps_per_rank": [round(p, 2) for p in per], "GBps_min": round(min(per), 2),
                      "GBps_aggregate": round(sum(per), 2),
                      "clips_per_s_ceiling_u8": round(world * 80 / (max(float(x.item()) for x in allms) * 1e-3), 1)}))
if world > 1:
    dist.destroy_process_group()
