#!/bin/bash
# Two-GPU session: NCCL tests, forward bench at N=2, training bench at N=2.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
echo "=== multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/pytest_multi.txt
echo "=== bench n2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n2.txt | cut -c1-400
echo "=== bench train n2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train_m224 --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_train_n2.txt | cut -c1-600
echo "=== eval driver n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m x3d_tf_b200.eval --cfg X3D_XS --model_folder /tmp --synthetic 12 --gpus 2 --allow_random_init 2>&1 | tail -3

echo "=== train driver n2"; rm -rf /tmp/x3d_train_n2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 -m x3d_tf_b200.train --config X3D_XS --model_dir /tmp/x3d_train_n2 --synthetic 16 --batch_size 4 --crop_size 64 --steps_per_epoch 2 --epochs 2 --num_gpus 2 2>&1 | tail -8; ls /tmp/x3d_train_n2
