#!/bin/bash
# Round capture: GPU tests, bench (both arms), ncu launch list and one --set full capture of K1 at full C4 size.
TAG=${1:-r01b}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_ref.log 2>&1; tail -1 gpurun_out/bench_${TAG}_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_identity2 -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/
