#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
for f in off first off first; do
  X3D_FUSE_EXPAND=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_fuse_$f.json 2> gpurun_out/r2_bench_fuse_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_fuse_$f.json").read().strip().splitlines()[-1])
kc=d["kernel_classes"]
print("fuse=$f", round(d["value"],1), round(d["ms_per_step"],3), {k:(v["ms"],v["launches"]) for k,v in kc.items() if isinstance(v,dict) and "ms" in v})
PY
done
