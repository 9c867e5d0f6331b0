#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/prof_dww.py 2>&1 | tee gpurun_out/prof_dww.txt
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_wgrad -s 1 -c 1 -f -o gpurun_out/dww_s4 python tools/prof_dww.py --reps 1 --only s4 > gpurun_out/ncu_dww.log 2>&1
tail -2 gpurun_out/ncu_dww.log
fi
