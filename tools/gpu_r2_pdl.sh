#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_pdl_pytest.txt
for f in 0 1 0 1; do
  X3D_PDL=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_pdl_$f.json 2> gpurun_out/r2_bench_pdl_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_pdl_$f.json").read().strip().splitlines()[-1])
print("pdl=$f", round(d["value"],1), round(d["ms_per_step"],3))
PY
done
