#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 2 -c 1 -f -o gpurun_out/stem_tc_v2 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_stem.log 2>&1
tail -2 gpurun_out/ncu_stem.log
