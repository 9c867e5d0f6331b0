#!/bin/bash
# ncu --set full of the two kernels changed most in the second half of round 2, final state:
# the planar channelwise kernel (stage-2 stride-1 layer) and the projection GEMM with the SE/swish prologue (stage 3)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw_planar -s 2 -c 1 \
    -o gpurun_out/r2_dw_planar_final -f python tools/prof_layers.py dwp --size 256 --clips 16 --reps 1 > gpurun_out/r2_ncu_dwp_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:pw_tc_kernelILb1 -s 3 -c 1 \
    -o gpurun_out/r2_pw_pro_final -f python tools/prof_layers.py pw --size 256 --clips 80 --reps 1 > gpurun_out/r2_ncu_pwpro_final.log 2>&1
tail -2 gpurun_out/r2_ncu_dwp_final.log gpurun_out/r2_ncu_pwpro_final.log
