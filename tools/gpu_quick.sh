#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_io.py -m gpu -q -p no:cacheprovider -x -k "stem" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_reference_golden.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.txt | cut -c1-300
