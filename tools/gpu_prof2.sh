#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/prof_layers.py pw --size 256 --clips 80 --reps 5 2>&1 | tee gpurun_out/prof_pw_256.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python tools/prof_train.py --clips 32 > gpurun_out/prof_train.log 2>&1
python tools/launch_summary.py gpurun_out/train_launches.csv | head -40
