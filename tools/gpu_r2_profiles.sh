#!/bin/bash
# Round-2 committed artefacts of the default bench command: launch list, DRAM traffic of the stencil
# kernel, one ncu --set full capture of it, then the bench lines themselves (not under a profiler).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_bench.csv 97 > gpurun_out/r2_launches_bench_summary.txt; head -16 gpurun_out/r2_launches_bench_summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:dw_tma|dw_planar" --csv --log-file gpurun_out/r2_dw_traffic.csv $B > gpurun_out/r2_dw_traffic.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/r2_clocks.csv &
SMI=$!
timeout 1200 python bench.py 2> gpurun_out/r2_bench_stderr.log | tail -1 > gpurun_out/r2_bench_m256x10_n1.json
kill $SMI
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_reference.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_m256x10_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['frac'],3), round(d['roofline']['fma_frac'],3), 'cpu', d['cpu_baseline'])
for c in d.get('configs',[]): print(c)
r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['config'])
PY
