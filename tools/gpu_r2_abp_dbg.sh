#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
(X3D_ABP_DEBUG=16 timeout 300 python tools/prof_layers.py ab2 --size 256 --clips 80 --reps 1 2>&1 | grep "abp plan" | sort -u
for na in 2 3 4; do
  echo "== X3D_ABP_NA=$na"
  X3D_ABP_NA=$na timeout 300 python tools/prof_layers.py ab2 --size 256 --clips 80 --reps 5 2>&1 | grep "ab2"
done) | tee gpurun_out/r2_abp_dbg.txt
