#!/bin/bash
# ncu --set full of the projection GEMM with the SE/swish prologue (pw_tc_kernel<true>): stage-3 and stage-5 shapes
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
# pro launches in prof_layers pw --reps 1: two per stage (warm-up + timed): 0,1 = s2; 2,3 = s3; 4,5 = s4; 6,7 = s5
for pick in 3 7; do
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pw_tc_kernelILb1 -s $pick -c 1 \
    -o gpurun_out/r2_pw_pro_$pick -f python tools/prof_layers.py pw --size 256 --clips 80 --reps 1 > gpurun_out/r2_ncu_pwpro_$pick.log 2>&1
tail -2 gpurun_out/r2_ncu_pwpro_$pick.log
done
