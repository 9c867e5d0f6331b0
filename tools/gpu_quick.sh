#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_io.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x -k "io or se_mlp or whole_model or eval_driver or evaluate" 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_q.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.txt').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['e2e_uint8']['value'])
print({k:(v['ms'] if isinstance(v,dict) and 'ms' in v else v) for k,v in d['kernel_classes'].items()})
PY
