#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_training.py tests/test_gpu_train_ops.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_pytest_train.txt
