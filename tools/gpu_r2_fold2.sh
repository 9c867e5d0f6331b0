#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_fold2_pytest.txt
for f in 0 1; do
  X3D_FOLD_SHORTCUT=$f timeout 600 python bench.py --workload m224 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_m224_fold_$f.json 2> gpurun_out/r2_bench_m224_fold_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_m224_fold_$f.json").read().strip().splitlines()[-1])
kc=d["kernel_classes"]
print("m224 fold=$f", round(d["value"],1), round(d["ms_per_step"],3), {k:(v["ms"],v["launches"]) for k,v in kc.items() if isinstance(v,dict) and "ms" in v})
PY
done
