#!/bin/bash
# Regenerates the committed ncu artefacts of the default bench command.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv 105 > gpurun_out/launches_bench_summary.txt; cat gpurun_out/launches_bench_summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dw_tma --csv --log-file gpurun_out/dw_traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/dw_traffic.log 2>&1
tail -3 gpurun_out/dw_traffic.csv
for w in s182 l356 m224 xs160; do timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$w.txt; python -c "
import json; d=json.loads(open('gpurun_out/bench_$w.txt').read()); print('$w', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['e2e']['value'])"; done
