#!/bin/bash
# racecheck on one case per TMA / mbarrier kernel, hazard records kept (racecheck has no model of
# mbarrier / async-proxy synchronisation: what it reports has to be read record by record)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for sel in "test_channelwise and 1-4-16-16-56-2 and dtype1" "test_fused_expand_channelwise_persistent and 2-4-16-16-24-56-1" "test_pointwise_tcgen05_plain"; do
  echo "=== -k '$sel'"
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 12 --show-backtrace no python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$sel" 2>&1 | grep -v "Host Frame" | head -120
done > gpurun_out/r2_sanitizer_racecheck_detail.txt 2>&1
grep -c "hazard" gpurun_out/r2_sanitizer_racecheck_detail.txt
grep -E "RACECHECK SUMMARY|passed|failed|===" gpurun_out/r2_sanitizer_racecheck_detail.txt
