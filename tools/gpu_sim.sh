#!/bin/bash
# similarity kernel: parity tests, throughput, optional ncu capture (C2 size)
TAG=${1:-sim}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "simil or smoke or golden or dropin" 2>&1 | tail -3
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_$TAG.log | cut -c1-420
if [ -n "$2" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_similarity2 -c 1 \
    -o gpurun_out/prof_$TAG -f python tools/bench_stats.py --only similarity --workloads C2 --repeats 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
fi
