#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for cfg in "256" "512"; do
X3D_PAIR_MAX_N=$cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_q.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.txt').read().strip().splitlines()[-1])
print('max_n=$cfg', round(d['value'],1), round(d['ms_per_step'],3), {k:(v['ms'] if isinstance(v,dict) and 'ms' in v else v) for k,v in d['kernel_classes'].items() if k in ('a','b','c','shortcut')})
PY
done
