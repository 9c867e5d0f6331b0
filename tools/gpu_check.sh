#!/bin/bash
# what the driver runs at round end: the GPU test suite, smoke(), the default bench line
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.txt
