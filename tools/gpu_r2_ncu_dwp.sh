#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw_planar -s ${NCU_SKIP:-2} -c 1 \
    -o gpurun_out/${NCU_OUT:-r2_dw_planar} -f python tools/prof_layers.py dwp --size 256 --clips 16 --reps 1 > gpurun_out/r2_ncu_dwp.log 2>&1
tail -3 gpurun_out/r2_ncu_dwp.log
