#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x -k "tcgen05 or tc_path or bf16 or full_size" 2>&1 | tail -4
timeout 600 python tools/prof_layers.py pw --size 256 --clips 80 --reps 5 2>&1
