#!/usr/bin/env python
"""Benchmark of the X3D forward path (BASELINE.json metric: X3D-M clips/sec at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W                      # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU reference arm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU

A "step" = one forward pass over one batch of synthetic clips of the workload's shape
(default workload: BASELINE.json configs[2], X3D-M 10-view eval at 16x256x256, bf16,
8 videos = 80 clips per GPU per step).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (variant, T, S, views, default clips/GPU/step, dtype)
    "xs160": ("X3D_XS", 4, 160, 1, 8, "float32"),          # BASELINE configs[0]
    "s182": ("X3D_S", 13, 182, 1, 64, "bfloat16"),          # configs[1]
    "m256x10": ("X3D_M", 16, 256, 10, 80, "bfloat16"),      # configs[2]  (metric config)
    "m224": ("X3D_M", 16, 224, 1, 64, "bfloat16"),          # north_star target shape
    "l356": ("X3D_L", 16, 356, 1, 32, "bfloat16"),          # configs[3]
    "train_m224": ("X3D_M", 16, 224, 1, 32, "float32"),     # configs[4]: one training step
}
METRIC = "X3D-M clips/sec"
FALLBACK_HBM_GBS, FALLBACK_BF16_TFLOPS = 6650.0, 1590.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return FALLBACK_HBM_GBS, FALLBACK_BF16_TFLOPS, "fallback"


def algorithmic_work(arch, T, H, W, esize):
    """Compulsory bytes and MACs per clip and per kernel class, TRUE channel counts
    (each kernel's logical input read once + output written
    once; weights excluded).  Classes follow the launches this build issues."""
    from x3d_tf_b200.arch import plan_shapes
    plan = plan_shapes(arch, T, H, W)
    c1 = arch.stem_channels
    out = {k: {"bytes": 0.0, "macs": 0.0} for k in
           ("stem", "a", "b", "ab", "c", "shortcut", "conv5", "head")}
    out["stem"]["bytes"] = (plan.input.P * 3 + plan.stem.P * c1) * esize
    out["stem"]["macs"] = plan.stem.P * c1 * (27 + arch.temp_filter)
    level = plan.stem
    for b in arch.blocks:
        p_in = level.P
        p_out = plan.stages[b.stage].P
        out["a"]["bytes"] += (b.cin + b.cinner) * p_in * esize
        out["a"]["macs"] += b.cin * b.cinner * p_in
        out["b"]["bytes"] += b.cinner * (p_in + p_out) * esize
        out["b"]["macs"] += 27 * b.cinner * p_out
        # fused expand+channelwise launch: only the block input and the channelwise output touch HBM
        out["ab"]["bytes"] += (b.cin * p_in + b.cinner * p_out) * esize
        out["ab"]["macs"] += b.cin * b.cinner * p_in + 27 * b.cinner * p_out
        out["c"]["bytes"] += (b.cinner + 2 * b.cout) * p_out * esize      # in + residual + out
        out["c"]["macs"] += b.cinner * b.cout * p_out
        if b.has_shortcut:
            out["shortcut"]["bytes"] += (b.cin + b.cout) * p_out * esize
            out["shortcut"]["macs"] += b.cin * b.cout * p_out
        level = plan.stages[b.stage]
    cl, c5 = arch.blocks[-1].cout, arch.conv5_channels
    out["conv5"]["bytes"] = (cl + c5) * level.P * esize
    out["conv5"]["macs"] = cl * c5 * level.P
    out["head"]["bytes"] = c5 * level.P * esize
    out["head"]["macs"] = c5 * arch.fc1_channels + arch.fc1_channels * arch.num_classes
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def build_model(workload, graph=True):
    from x3d_tf_b200 import model as M
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.synth import synthetic_weights
    variant, T, S, views, clips, dtype = WORKLOADS[workload]
    cfg = get_config(variant, freeze=False)
    cfg.TEST.NUM_TEMPORAL_VIEWS, cfg.TEST.NUM_SPATIAL_CROPS = views, 1
    cfg.freeze()
    M.reset_block_counters()
    m = M.X3D(cfg, dtype=dtype, use_cuda_graph=graph)
    arch = build_arch(cfg)
    W = synthetic_weights(arch, seed=1111)
    m.set_weights_dict(W)
    return m, cfg, arch, W


def device_clips(n, T, S, cfg, dtype, device, seed):
    """Synthetic normalised clips generated on the device (uniform u8 pixels -> (x/255-mean)/std)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u8 = torch.randint(0, 256, (n, T, S, S, 3), generator=g, device=device, dtype=torch.uint8)
    mean = torch.tensor(cfg.DATA.MEAN, device=device, dtype=torch.float32)
    std = torch.tensor(cfg.DATA.STD, device=device, dtype=torch.float32)
    out = torch.empty((n, T, S, S, 3), device=device, dtype=dtype)
    for i in range(n):          # clip by clip: bounded temporary memory
        out[i] = ((u8[i].float() / 255.0 - mean) / std).to(dtype)
    return out


def cpu_reference_clips_per_s(workload, steps, warmup, budget_s, clips_per_step=None):
    """The reference's algorithm on the host cores: the torch-CPU fp32 restatement in oracle/
    (TensorFlow itself cannot be installed here -- DESIGN.md).  Returns (clips/s, info)."""
    from oracle import x3d_oracle as O
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.synth import synthetic_clips, synthetic_weights
    variant, T, S, views, _, _ = WORKLOADS[workload]
    cfg = get_config(variant)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = O.OracleSpec.from_cfg(cfg, )
    spec.num_preds = 1
    W = synthetic_weights(build_arch(cfg), seed=1111)
    one = synthetic_clips(1, T, S, S, cfg.DATA.MEAN, cfg.DATA.STD, seed=1)
    t0 = time.perf_counter()
    O.forward(W, spec, one, torch.float32)
    t1 = time.perf_counter() - t0            # includes first-touch cost; upper bound per clip
    if clips_per_step is None:
        clips_per_step = int(max(1, min(8, budget_s / max(t1, 1e-3) / max(steps + warmup, 1))))
    x = synthetic_clips(clips_per_step, T, S, S, cfg.DATA.MEAN, cfg.DATA.STD, seed=2)
    for _ in range(warmup):
        O.forward(W, spec, x, torch.float32)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward(W, spec, x, torch.float32)
    dt = time.perf_counter() - t0
    return clips_per_step * steps / dt, {
        "cores": cores, "kind": "port",
        "sample": f"{steps} steps x {clips_per_step} clips of {variant} {T}x{S}x{S} fp32 "
                  f"(torch-CPU restatement of model.py, {cores} threads), {warmup} warm-up",
        "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    variant, T, S, views, clips, dtype = WORKLOADS[args.workload]
    v, info = cpu_reference_clips_per_s(args.workload, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{variant} {T}x{S}x{S}, {views}-view eval clips",
                       "note": "CPU restatement of the reference (TensorFlow unavailable)"},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": info["cores"],
                             "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch.distributed as dist
    from x3d_tf_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    variant, T, S, views, clips, dtype_name = WORKLOADS[args.workload]
    if args.clips:
        clips = args.clips
    clips -= clips % views
    tdt = torch.bfloat16 if dtype_name == "bfloat16" else torch.float32
    esize = 2 if tdt == torch.bfloat16 else 4
    model, cfg, arch, _ = build_model(args.workload, graph=True)
    x = device_clips(clips, T, S, cfg, tdt, device, seed=1111 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: graph replay on the captured input buffer
    model(x)                                   # eager warm-up + capture
    static_in = model.static_input(x.shape, tdt)
    static_in.copy_(x)
    del x
    for _ in range(max(args.warmup, 3)):
        probs = model(static_in)
    barrier()
    ops.Profiler.launches = 0
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        probs = model(static_in)
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = clips * world * args.steps / (ms_max / 1e3)

    # ---- per-class attribution: one eager, event-bracketed pass over the same batch
    model._use_graph = False
    ops.Profiler.start()
    model(static_in)
    recs = ops.Profiler.stop()
    launches_per_step = len(recs)
    model._use_graph = True
    work = algorithmic_work(arch, T, S, S, esize)
    hbm, tflops, peak_kind = peaks()
    classes = {}
    for tag, name, dt_ms in recs:
        if tag.startswith("head_"):
            sub = classes.setdefault("head_parts", {})
            sub[tag] = round(sub.get(tag, 0.0) + dt_ms, 4)
            tag = "head"
        c = classes.setdefault(tag, {"ms": 0.0, "launches": 0})
        c["ms"] += dt_ms
        c["launches"] += 1
    for tag, c in classes.items():
        if tag == "head_parts":
            continue
        if tag in work:
            c["GBps"] = work[tag]["bytes"] * clips / (c["ms"] * 1e-3) / 1e9
            c["hbm_frac"] = c["GBps"] / hbm
            c["tflops"] = 2 * work[tag]["macs"] * clips / (c["ms"] * 1e-3) / 1e12
        c["ms"] = round(c["ms"], 4)
    fused = "ab" in classes and "b" not in classes
    if fused:
        b = classes["ab"]
        b_bytes = work["ab"]["bytes"] * clips
        kname = "ab_fused_kernel (expand 1x1x1 + BN + ReLU -> channelwise 3x3x3 + BN + SE sums)"
    else:
        b = classes.get("b", {"ms": float("nan"), "launches": 0})
        b_bytes = work["b"]["bytes"] * clips
        kname = "dw_tma_kernel (channelwise 3x3x3 + BN + SE sums)"
    achieved = b_bytes / (b["ms"] * 1e-3) / 1e9
    # DRAM bytes per launch of the same kernel from the committed ncu capture (read + write,
    # averaged over the launches of one step, like `achieved`); null if no capture matches
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_dw_traffic.json")
    if not fused and os.path.exists(tp):
        with open(tp) as f:
            td = json.load(f)
        if td.get("workload") == args.workload and td.get("clips") == clips and \
                kname.startswith(td.get("kernel", "dw_tma_kernel")):
            traffic = td["bytes_per_launch"]
    roofline = {"kernel": kname, "bound": "hbm",
                "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "algorithmic_bytes_per_launch": b_bytes / max(b["launches"], 1),
                "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                "launches_per_step": b["launches"],
                "algorithmic_bytes_per_step": b_bytes,
                "avg_launch_ms": b["ms"] / max(b["launches"], 1),
                "how": "CUDA events around each launch of one eager pass over the timed batch"}
    if fused:
        # the stencil half runs on the fp32 FMA pipe: 27 MAC per output against the packed-FFMA2
        # rate measured for this operand pattern (profiles/r01_microbench_fma_copy.txt)
        roofline["stencil_fma_tmacs"] = work["b"]["macs"] * clips / (b["ms"] * 1e-3) / 1e12
        roofline["stencil_fma_frac_of_31.4_TFMA/s"] = roofline["stencil_fma_tmacs"] / 31.4
        roofline["unfused_equivalent_GBps"] = (work["a"]["bytes"] + work["b"]["bytes"]) * clips / (b["ms"] * 1e-3) / 1e9
    total_bytes = sum(w["bytes"] for k, w in work.items() if k != "ab") * clips
    pw_macs = (work["a"]["macs"] + work["c"]["macs"]) * clips
    pw_ms = classes.get("a", classes.get("ab", {"ms": 0}))["ms"] + classes.get("c", {"ms": 0})["ms"]
    extra = {"whole_model_hbm_frac_layerwise": total_bytes / (ms_max / args.steps * 1e-3) / 1e9 / hbm,
             "pointwise_tensor_pipe_util": (2 * pw_macs / (pw_ms * 1e-3) / 1e12 / tflops) if pw_ms else None,
             "kernel_classes": classes}

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the region
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty((clips, T, S, S, 3), dtype=tdt).pin_memory()
        host_in.copy_(static_in.cpu())
        want = probs.float().cpu()

        def e2e_run(n):
            # public API: X3D.predict(iterable of HOST batches) -> host probabilities; every step
            # copies its clips host->device (pinned, copy stream, overlapped with the previous
            # step's forward) and its probabilities device->host inside the timed region
            outs = None
            for outs in model.predict(host_in for _ in range(n)):
                pass
            return outs

        got = e2e_run(3)
        if not torch.allclose(got, want, rtol=0, atol=1e-6):
            raise SystemExit("bench.py: predict() result differs from the device-resident result")
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_run(n_e2e)
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": clips * world * n_e2e / (float(te.item()) / 1e3), "unit": "clips/s",
               "h2d_bytes_per_step": host_in.numel() * host_in.element_size(),
               "d2h_bytes_per_step": want.numel() * 4, "steps": n_e2e,
               "api": "X3D.predict on pinned-host clips: per step H2D copy (copy stream, overlapped "
                      "with the previous step's forward), forward, D2H of the probabilities"}

    # ---- the same, fed with decoded uint8 frames: the input stage (utils.normalize) runs on the
    # device inside the stem's loader, so the host->device copy is half the bf16 one
    e2e_u8 = None
    if not args.no_e2e and tdt == torch.bfloat16:
        g8 = torch.Generator()
        g8.manual_seed(2222 + rank)
        host_u8 = torch.randint(0, 256, (clips, T, S, S, 3), dtype=torch.uint8, generator=g8).pin_memory()
        want8 = model(host_u8.to(device)).float().cpu()
        got8 = None
        for got8 in model.predict(host_u8 for _ in range(3)):
            pass
        if not torch.allclose(got8, want8, rtol=0, atol=1e-6):
            raise SystemExit("bench.py: predict(uint8) differs from the device-resident uint8 result")
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in model.predict(host_u8 for _ in range(n_e2e)):
            pass
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_u8 = {"value": clips * world * n_e2e / (float(te.item()) / 1e3), "unit": "clips/s",
                  "h2d_bytes_per_step": host_u8.numel(), "d2h_bytes_per_step": want8.numel() * 4,
                  "steps": n_e2e,
                  "api": "X3D.predict on pinned-host uint8 frames (utils.normalize fused into the stem loader)"}
        extra["e2e_uint8"] = e2e_u8

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, info = cpu_reference_clips_per_s(args.workload, steps=2, warmup=1, budget_s=20.0)
            cpu = {"value": v, "unit": "clips/s", "cores": info["cores"], "kind": info["kind"],
                   "sample": info["sample"]}
        line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if tdt == torch.bfloat16 else "f32",
                "data": "synthetic",
                "config": {"workload": f"{variant} {T}x{S}x{S}, {views}-view eval, "
                                       f"{clips // views} videos = {clips} clips per GPU per step",
                           "weights": "synthetic (checkpoint names/shapes; data shards absent)",
                           "l2": f"inputs larger than L2: the clip batch is "
                                 f"{clips * T * S * S * 3 * esize / 1e6:.0f} MB and every "
                                 "intermediate tensor is larger",
                           "cuda_graph": True, "parallelism": f"dp{world} (videos sharded, no collective)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline, "cpu_baseline": cpu, **extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE configs[4]: one X3D-M training step (forward, backward, NCCL gradient all-reduce,
    SGD-Nesterov) per bench step; batch 32 per GPU, 16x224x224, fp32 (the reference's default
    precision; train.py:85-152).  value = clips trained per second over all ranks."""
    import torch.distributed as dist
    from x3d_tf_b200 import _lib
    from x3d_tf_b200.arch import build_arch
    from x3d_tf_b200.config import get_config
    from x3d_tf_b200.synth import synthetic_weights
    from x3d_tf_b200.training import X3DTrainer, lr_schedule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    variant, T, S, _, clips, _ = WORKLOADS[args.workload]
    if args.clips:
        clips = args.clips
    cfg = get_config(variant)
    arch = build_arch(cfg)
    tr = X3DTrainer(cfg, device=device, world=world).load(synthetic_weights(arch, seed=1111))
    x = device_clips(clips, T, S, cfg, torch.float32, device, seed=1111 + rank)
    g = torch.Generator(device=device)
    g.manual_seed(7 + rank)
    labels = torch.randint(0, cfg.NETWORK.NUM_CLASSES, (clips,), generator=g, device=device, dtype=torch.int32)
    lr = lr_schedule(cfg, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        loss = tr.step(x, labels, lr)
    barrier()
    c0 = _lib.calls
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(x, labels, lr)
    e1.record()
    barrier()
    clocks = sampler.finish()
    launches = _lib.calls - c0
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = clips * world * args.steps / (ms_max / 1e3)

    # end to end: clips and labels come from pinned host memory every step, the loss goes back
    host_x = torch.empty(x.shape, dtype=torch.float32).pin_memory()
    host_x.copy_(x.cpu())
    host_l = labels.cpu().pin_memory()
    n_e2e = max(3, min(args.steps, 5))
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_e2e):
        x.copy_(host_x, non_blocking=True)
        labels.copy_(host_l, non_blocking=True)
        host_loss = tr.step(x, labels, lr).float().cpu()
    s1.record()
    barrier()
    te = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    hbm, _, peak_kind = peaks()
    work = algorithmic_work(arch, T, S, S, 4)
    fwd_bytes = sum(w["bytes"] for k, w in work.items() if k != "ab") * clips
    if rank == 0:
        line = {"metric": "X3D-M training clips/sec", "value": value, "unit": "clips/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{variant} training step {T}x{S}x{S}, batch {clips} per GPU, SGD-Nesterov "
                                       f"momentum {tr.momentum}, L2, dropout {tr.dropout}, batch-statistics BN",
                           "exchange": "one NCCL sum all-reduce of the flat fp32 gradient arena "
                                       f"({tr.layout.size * 4 / 1e6:.2f} MB)" if world > 1 else "single rank: no exchange",
                           "l2": f"inputs larger than L2: the clip batch is {x.numel() * 4 / 1e6:.0f} MB",
                           "parallelism": f"dp{world}"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": clips * world * n_e2e / (float(te.item()) / 1e3), "unit": "clips/s",
                        "h2d_bytes_per_step": host_x.numel() * 4 + host_l.numel() * 4,
                        "d2h_bytes_per_step": host_loss.numel() * 4, "steps": n_e2e,
                        "api": "X3DTrainer.step on clips/labels copied from pinned host memory, per-clip losses read back"},
                "roofline": {"kernel": "whole training step", "bound": "hbm",
                             "achieved": 3 * fwd_bytes / (ms_max / args.steps * 1e-3) / 1e9, "peak": hbm,
                             "unit": "GB/s", "frac": 3 * fwd_bytes / (ms_max / args.steps * 1e-3) / 1e9 / hbm,
                             "traffic": None, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                             "how": "3 x the forward pass's algorithmic fp32 bytes (SURVEY.md 8d estimate) / step time"},
                "cpu_baseline": None, "loss_mean": float(loss.float().mean().item())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="m256x10", choices=list(WORKLOADS))
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU per step (override)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("train"):
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
