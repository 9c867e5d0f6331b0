#!/bin/bash
# ncu --set full of the persistent fused kernel on the stage-2 stride-1 layer (16 clips of 16x256^2)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_persist -s ${NCU_SKIP:-2} -c 1 \
    -o gpurun_out/${NCU_OUT:-r2_ab_persist} -f python tools/prof_layers.py ab2 --size 256 --clips 16 --reps 1 > gpurun_out/r2_ncu_ab2.log 2>&1
tail -3 gpurun_out/r2_ncu_ab2.log
