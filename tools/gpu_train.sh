#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x -k "not tcgen05 and not tc_path and not bf16 and not full_size" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python tools/prof_train.py --clips 32 > gpurun_out/prof_train.log 2>&1
python tools/launch_summary.py gpurun_out/train_launches.csv > gpurun_out/train_launch_summary.txt; head -14 gpurun_out/train_launch_summary.txt
timeout 900 python bench.py --workload train_m224 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_train.txt | cut -c1-330
