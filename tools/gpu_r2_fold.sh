#!/bin/bash
# shortcut conv folded into the projection GEMM: parity, then bench A/B
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_reference_golden.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_fold_pytest.txt
for f in 0 1; do
  X3D_FOLD_SHORTCUT=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_fold_$f.json 2> gpurun_out/r2_bench_fold_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_fold_$f.json").read().strip().splitlines()[-1])
kc=d["kernel_classes"]
print("fold=$f", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:(v["ms"],v["launches"]) for k,v in kc.items() if isinstance(v,dict) and "ms" in v})
PY
done
