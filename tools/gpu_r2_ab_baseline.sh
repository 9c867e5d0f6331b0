#!/bin/bash
# round 2, call 1: per-layer times of the fused expand+channelwise kernel next to the unfused pair,
# and one ncu --set full capture of the fused kernel on the stage-2 stride-1 layer.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 600 python tools/prof_layers.py ab --size 256 --clips 80 --reps 5 2>&1 | tee gpurun_out/r2_prof_ab_256.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_fused -s 2 -c 1 \
    -o gpurun_out/r2_ab_fused_v1 python tools/prof_layers.py ab --size 256 --clips 16 --reps 1 > gpurun_out/r2_ncu_ab.log 2>&1
tail -3 gpurun_out/r2_ncu_ab.log
