#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_reference_golden.py tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_head_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r2_bench_head.json 2> gpurun_out/r2_bench_head.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_head.json").read().strip().splitlines()[-1])
kc=d["kernel_classes"]
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:(v["ms"],v["launches"]) for k,v in kc.items() if isinstance(v,dict) and "ms" in v}, kc.get("head_parts"))
PY
