#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_reference_golden.py tests/test_gpu_io.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_q.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.txt').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step']}, d['e2e']['value'], d['e2e_uint8']['value'], d['roofline']['frac'])
print({k:(v['ms'] if isinstance(v,dict) and 'ms' in v else v) for k,v in d['kernel_classes'].items()})
PY
