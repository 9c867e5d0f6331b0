#!/bin/bash
# compute-sanitizer over the op-level GPU tests (small shapes): memcheck on all of them, racecheck and
# synccheck on the TMA / mbarrier / tcgen05 kernels
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
SEL='persistent or channelwise or tcgen05 or stem or se_mlp or head_fc or avgpool or gather'
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_ops.py -x -q -m gpu -k "$SEL" 2>&1 | tail -25 ) > gpurun_out/r2_sanitizer_memcheck.txt
echo "memcheck exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_sanitizer_memcheck.txt
( timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "persistent or (channelwise and not fused) or pointwise_tcgen05_plain or stem_tcgen05" 2>&1 | tail -40 ) > gpurun_out/r2_sanitizer_racecheck.txt
( timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "persistent or pointwise_tcgen05_plain" 2>&1 | tail -25 ) > gpurun_out/r2_sanitizer_synccheck.txt
tail -6 gpurun_out/r2_sanitizer_memcheck.txt gpurun_out/r2_sanitizer_racecheck.txt gpurun_out/r2_sanitizer_synccheck.txt
