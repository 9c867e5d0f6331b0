#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "channelwise" 2>&1 | tail -4
timeout 300 python tools/prof_layers.py dwp --size 256 --clips 120 --reps 5 2>&1 | grep -A1 "s5 8x8"
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -3
for f in tma auto; do
  X3D_CHANNELWISE=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_mc4_$f.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_mc4_$f.json").read().strip().splitlines()[-1])
print("cw=$f", round(d["value"],1), round(d["ms_per_step"],3), d["kernel_classes"]["b"])
PY
done
