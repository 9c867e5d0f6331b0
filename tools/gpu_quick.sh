#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
X3D_PAIR_PIXELS=4 X3D_PAIR_ALIGNED=1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_reference_golden.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
for cfg in "1 0 256" "2 0 256" "4 0 256" "2 1 256" "4 1 256" "4 1 128" "2 0 256"; do
set -- $cfg
X3D_PAIR_PIXELS=$1 X3D_PAIR_ALIGNED=$2 X3D_PAIR_MAX_K=$3 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_q.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.txt').read().strip().splitlines()[-1])
print('pair=$1 aligned=$2 maxk=$3', round(d['value'],1), round(d['ms_per_step'],3), {k:(v['ms'] if isinstance(v,dict) and 'ms' in v else v) for k,v in d['kernel_classes'].items() if k in ('a','b','c','shortcut','se')})
PY
done
