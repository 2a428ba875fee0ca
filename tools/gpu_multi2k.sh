#!/bin/bash
# 2-GPU check after the K1 tile-order / K4 changes: multi-rank bit-identity + multi-GPU tests + bench at N=2
TAG=${1:-r01k}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tools/multigpu_check.py > gpurun_out/multigpu_${TAG}_n2.log 2>&1; echo "multigpu_check rc=$?"; grep '^{' gpurun_out/multigpu_${TAG}_n2.log | cut -c1-330
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -2
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 ) > gpurun_out/bench_${TAG}_n2.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_n2.log | tail -1 | cut -c1-600
