#!/bin/bash
# whole-model checks with the fused kernel in the default path + a quick bench
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_reference_golden.py tests/test_gpu_model.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2_pytest_model.txt
for mode in auto off all; do
X3D_FUSE_EXPAND=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs 2>&1 | tail -1 > gpurun_out/r2_bench_$mode.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode', round(d['value'],1), round(d['ms_per_step'],3), {k:(v['ms'],v['launches']) for k,v in d['kernel_classes'].items() if isinstance(v,dict) and 'ms' in v})
PY
done
