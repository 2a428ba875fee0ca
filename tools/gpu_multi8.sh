#!/bin/bash
# 8-GPU capture: multi-rank bit-identity check, bench.py at N=8 and N=4
TAG=${1:-r01j}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    tools/multigpu_check.py --shape 700x900,1500x333 > gpurun_out/multigpu_${TAG}_n8.log 2>&1; echo "multigpu_check rc=$?"; grep '^{' gpurun_out/multigpu_${TAG}_n8.log | cut -c1-330
for g in 8 4; do
    ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2954$g \
        bench.py --gpus $g ) > gpurun_out/bench_${TAG}_n$g.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_n$g.log | tail -1 | cut -c1-1700
done
