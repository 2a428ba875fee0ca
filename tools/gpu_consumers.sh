#!/bin/bash
# identity-matrix consumers: GPU parity tests, timings, optional ncu
TAG=${1:-cons}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_consumers_gpu.py -m gpu -x -q ) > gpurun_out/pytest_$TAG.log 2>&1; tail -15 gpurun_out/pytest_$TAG.log
timeout 600 python tools/bench_stats.py --only consumers --workloads C3,C4 --repeats 3 | tee gpurun_out/stats_$TAG.log
if [ -n "$2" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_identity_bits|k_row_stats|k_mis' -c 8 \
    -o gpurun_out/prof_$TAG -f python tools/bench_stats.py --only consumers --workloads C4 --repeats 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
fi
