#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -k "channelwise and not fused" -q -p no:cacheprovider -x 2>&1 | tail -5
for impl in v3 pw; do
  echo "== $impl"
  X3D_DW_IMPL=$impl timeout 600 python tools/prof_layers.py dw --size 256 --clips 80 --reps 5 --se 1 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw -s 2 -c 1 \
      -o gpurun_out/dw_pw_se1 -f python tools/prof_layers.py dw --size 256 --clips 16 --reps 1 --se 1 > gpurun_out/ncu_dw_pw_se1.log 2>&1
