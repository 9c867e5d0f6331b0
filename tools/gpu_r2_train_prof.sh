#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_train_launches.csv python tools/prof_train.py --clips 32 > gpurun_out/r2_prof_train.log 2>&1
python tools/launch_summary.py gpurun_out/r2_train_launches.csv | head -45 | tee gpurun_out/r2_train_launch_summary.txt
