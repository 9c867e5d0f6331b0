#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): per-launch key metrics, warp-stall
breakdown and the executed-instruction mix from the source page.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source] [--launch N]"""
import csv, io, subprocess, sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, rows[2:]


def main():
    path = sys.argv[1]
    hdr, rows = raw(path)
    for i, r in enumerate(rows):
        d = dict(zip(hdr, r))
        print(f"=== launch {i}: {d.get('Kernel Name', '')[:110]}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:75s} {d[k]}")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")),
                        reverse=True)
        if not stalls:
            stalls = sorted(((float(v), k) for k, v in d.items()
                             if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct") and v not in ("", "n/a")), reverse=True)
        for v, k in stalls[:8]:
            print(f"  stall {k.split('stalled_')[1][:40]:42s} {v:8.2f}")
    if "--source" in sys.argv:
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[1]
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        ops, samp, tot, stot = Counter(), Counter(), 0, 0
        for r in rows[2:]:
            if r and r[0] == "Kernel Name":
                break
            if len(r) < len(hdr) or r[0] == "Address":
                continue
            try:
                n, s = int(r[ie]), int(r[isamp])
            except ValueError:
                continue
            toks = r[ia].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            ops[op] += n; samp[op] += s; tot += n; stot += s
        print(f"--- instruction mix of launch 0: {tot} warp instructions, {stot} samples")
        for op, n in ops.most_common(18):
            print(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * samp[op] / max(stot, 1):5.1f}%")


if __name__ == "__main__":
    main()
