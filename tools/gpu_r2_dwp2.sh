#!/bin/bash
# planar channelwise kernel: op parity, whole-model fixtures, layer times at the three clip sizes, bench A/B
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_reference_golden.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_dwp2_pytest.txt
for sz in 224 182; do
  timeout 300 python tools/prof_layers.py dwp --size $sz --clips 80 --reps 5 2>&1 | tee gpurun_out/r2_prof_dwp_$sz.txt
done
for cw in tma auto; do
  X3D_CHANNELWISE=$cw timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_cw_$cw.json 2> gpurun_out/r2_bench_cw_$cw.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_cw_$cw.json").read().strip().splitlines()[-1])
print("$cw", d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("roofline"))
PY
done
