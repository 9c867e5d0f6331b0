#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_pytest_train2.txt
timeout 600 python bench.py --workload train_m224 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_train2.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_train2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms/step launches', d['gpu_launches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_train_launches.csv python tools/prof_train.py --clips 32 > gpurun_out/r2_prof_train.log 2>&1
python tools/launch_summary.py gpurun_out/r2_train_launches.csv | head -24 | tee gpurun_out/r2_train_launch_summary.txt
