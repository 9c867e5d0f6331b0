#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15
