#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
bash tools/gpu_r2_sanitizer2.sh
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_io.py -x -q -m gpu -k "channelwise or head_fc or eval_views" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs 2>&1 | tail -1 > gpurun_out/r2_bench_quick.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_quick.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],3), {k:(v['ms'],v['launches']) for k,v in d['kernel_classes'].items() if isinstance(v,dict) and 'ms' in v}, d['kernel_classes'].get('head_parts'))
PY
