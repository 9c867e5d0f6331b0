#!/bin/bash
# round 2: parity of the persistent fused expand+channelwise kernel, then per-layer times
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "persistent" 2>&1 | tail -15 | tee gpurun_out/r2_ab2_pytest.txt
timeout 600 python tools/prof_layers.py ab2 --size 256 --clips 80 --reps 5 2>&1 | grep ab2 | tee gpurun_out/r2_prof_ab2_256.txt
for d in 1 2; do
  echo "== X3D_ABP_DEBUG=$d"
  X3D_ABP_DEBUG=$d timeout 300 python tools/prof_layers.py ab2 --size 256 --clips 80 --reps 5 2>&1 | grep "ab2"
done | tee gpurun_out/r2_abp_dbg.txt
