#!/bin/bash
# fp32 (3xTF32) pointwise GEMM: per-shape CUDA-event times and one ncu --set full capture.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_pw32.py 2>&1 | tee gpurun_out/prof_pw32.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tf32x3 -s 2 -c 1 -f -o gpurun_out/pw32_s4_expand python tools/prof_pw32.py --reps 1 --only "s4 expand" > gpurun_out/ncu_pw32.log 2>&1
tail -2 gpurun_out/ncu_pw32.log
