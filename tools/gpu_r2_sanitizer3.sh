#!/bin/bash
# compute-sanitizer over the kernels added / changed in the second half of round 2: memcheck on the planar
# channelwise kernel, the pointwise kernel's second source / column means / two-group prologue, the cluster
# head GEMM and the stem; synccheck on the planar kernel and the cluster GEMM
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
SEL='channelwise_planar or pointwise_tcgen05 or head_fc or stem or colreduce or bn_'
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -x -q -m gpu -k "$SEL" 2>&1 | tail -12 ) > gpurun_out/r2_sanitizer_memcheck_b.txt
( timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "channelwise_planar or head_fc or shortcut_as_extra_k or column_means" 2>&1 | tail -8 ) > gpurun_out/r2_sanitizer_synccheck_b.txt
tail -5 gpurun_out/r2_sanitizer_memcheck_b.txt gpurun_out/r2_sanitizer_synccheck_b.txt
