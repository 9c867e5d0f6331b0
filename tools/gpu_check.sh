#!/bin/bash
# One GPU-box session: parity tests (separate processes so a CUDA fault in one group cannot
# poison the others), smoke, a short bench run.  Output -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1500 python -m pytest "$@" -q -p no:cacheprovider -x 2>&1 | tail -25 | tee gpurun_out/pytest_$name.txt; }
run ops_simt tests/test_gpu_ops.py -m gpu -k "not tcgen05 and not tc_path"
run ops_tc tests/test_gpu_ops.py -m gpu -k "tcgen05 or tc_path"
run io tests/test_gpu_io.py -m gpu
run model_fp32_simt tests/test_gpu_model.py -m gpu -k "fp32 or simt or standalone or checkpoint"
run model_tc tests/test_gpu_model.py -m gpu -k "not (fp32 or simt or standalone or checkpoint)"
run golden tests/test_reference_golden.py -m gpu
run train tests/test_gpu_train_ops.py tests/test_gpu_training.py -m gpu
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} 2>&1 | tail -5 | tee gpurun_out/bench.txt
echo "=== bench train"; timeout 900 python bench.py --workload train_m224 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_train.txt
