#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_reference_golden.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_q7_pytest.txt
timeout 300 python tools/prof_layers.py dwp --size 224 --clips 80 --reps 5 2>&1 | grep -A1 stride1 | tee gpurun_out/r2_prof_dwp_224_q7.txt
for f in tma auto; do
  X3D_CHANNELWISE=$f timeout 600 python bench.py --workload m224 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_m224_cw_$f.json 2> gpurun_out/r2_bench_m224_cw_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_m224_cw_$f.json").read().strip().splitlines()[-1])
kc=d["kernel_classes"]
print("m224 cw=$f", round(d["value"],1), round(d["ms_per_step"],3), {k:(v["ms"],v["launches"]) for k,v in kc.items() if isinstance(v,dict) and "ms" in v and k in "abc"}, round(d["roofline"]["frac"],3), round(d["roofline"]["fma_frac"],3))
PY
done
