#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_io.py -m gpu -q -p no:cacheprovider -x -k "stem" 2>&1 | tail -4
timeout 900 python tools/prof_layers.py stem --size 256 --clips 80 --reps 5 2>&1
