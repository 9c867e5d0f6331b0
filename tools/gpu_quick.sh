#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
# launches of prof_layers pw: per stage: a (warm + reps), c-se (warm+reps), c-plain (warm+reps); reps=1 -> 2 each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -s 1 -c 1 -o gpurun_out/pw_a_s2 -f python tools/prof_layers.py pw --size 256 --clips 40 --reps 1 > gpurun_out/ncu_pw_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_tc -s 5 -c 1 -o gpurun_out/pw_cplain_s2 -f python tools/prof_layers.py pw --size 256 --clips 40 --reps 1 > gpurun_out/ncu_pw_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
