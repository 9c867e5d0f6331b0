#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "channelwise_planar" 2>&1 | tail -12 | tee gpurun_out/r2_dwp_pytest.txt
timeout 600 python tools/prof_layers.py dwp --size 256 --clips 80 --reps 5 2>&1 | tee gpurun_out/r2_prof_dwp_256.txt
