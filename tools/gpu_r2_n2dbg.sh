#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs --no-e2e 2>/dev/null | tail -1 > gpurun_out/n2dbg.json
  python - "$*" <<PY
import json, sys
d=json.loads(open("gpurun_out/n2dbg.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"))
PY
}
run X3D_PDL=1
run X3D_PDL=0
run X3D_CHANNELWISE=tma
run X3D_PDL=0 X3D_CHANNELWISE=tma
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
