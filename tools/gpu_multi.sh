#!/bin/bash
# N-GPU capture (run with `gpurun --gpus N -- bash tools/gpu_multi.sh N tag`): GPU tests (incl. the
# multi-rank one), the *_all bit-identity check with the C3-sized timing, and bench.py at 1..N GPUs.
N=${1:-2}
TAG=${2:-r01c}
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/multigpu_check.py --big > gpurun_out/multigpu_${TAG}_n$N.log 2>&1; echo "multigpu_check rc=$?"; grep '^{' gpurun_out/multigpu_${TAG}_n$N.log | cut -c1-400
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-300
for g in 2 4 8; do
  if [ $g -le $N ]; then
    ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2954$g \
        bench.py --gpus $g ) > gpurun_out/bench_${TAG}_n$g.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_n$g.log | tail -1 | cut -c1-300
  fi
done
