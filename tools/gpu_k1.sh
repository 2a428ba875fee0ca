#!/bin/bash
# K1 change check: parity tests, bench, ncu --set full of one k_identity2 launch at full C4 size
TAG=${1:-k1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=150 2>&1 | grep -v Warning | tail -6
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_identity2 -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
