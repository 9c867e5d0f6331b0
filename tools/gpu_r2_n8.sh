#!/bin/bash
# 8-GPU session: concurrent pinned-host -> device copy bandwidth (the e2e ceiling) and the bench line at N = 8
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
N=${N:-8}
for n in 1 2 4 $N; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/copy_bench.py 2>/dev/null | tail -1
done | tee gpurun_out/r2_copy_bench.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_m256x10_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_m256x10_n$N.json').read().strip().splitlines()[-1])
print('N=$N value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['e2e']['h2d_bytes_per_step'], 'bf16 e2e', round(d.get('e2e_bf16_clips',{}).get('value',0),1))
for c in d.get('configs',[]): print(c['workload'], round(c['value'],1), round(c['ms_per_step'],2), c.get('allreduce_ms'))
PY
