#!/bin/bash
# N-GPU capture: multi-rank bit-identity check (incl. representatives_all) and bench.py at 1..N GPUs, both arms at N=1
N=${1:-2}
TAG=${2:-r01f}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/multigpu_check.py > gpurun_out/multigpu_${TAG}_n$N.log 2>&1; echo "multigpu_check rc=$?"; grep '^{' gpurun_out/multigpu_${TAG}_n$N.log | cut -c1-400
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-2500
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_ref.log 2>&1; tail -1 gpurun_out/bench_${TAG}_ref.log | cut -c1-600
for g in 2 4 8; do
  if [ $g -le $N ]; then
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2954$g \
        bench.py --gpus $g ) > gpurun_out/bench_${TAG}_n$g.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_n$g.log | tail -1 | cut -c1-1800
  fi
done
