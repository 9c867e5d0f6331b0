#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -x -q -m gpu -k "tf32x3_tcgen05" 2>&1 | tail -15 | tee gpurun_out/r2_pytest_tf32.txt
timeout 1500 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_pytest_train.txt
for g in mma_sync tcgen05; do
X3D_TRAIN_GEMM=$g timeout 600 python bench.py --workload train_m224 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_train_$g.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_train_$g.json').read().strip().splitlines()[-1])
print('$g', round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms/step launches', d['gpu_launches'])
PY
done
