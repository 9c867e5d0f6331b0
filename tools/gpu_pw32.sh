#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
echo "== pipe"; timeout 300 python tools/prof_pw32.py 2>&1 | tee gpurun_out/prof_pw32.txt
echo "== no pipe"; X3D_PW32_NOPIPE=1 timeout 300 python tools/prof_pw32.py 2>&1 | tee gpurun_out/prof_pw32_nopipe.txt
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
for v in "" 1; do
X3D_PW32_NOPIPE=$v timeout 900 python bench.py --workload train_m224 --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-250
done
