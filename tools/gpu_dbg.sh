#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "simil or smoke or golden or dropin or pytrimal" 2>&1 | tail -3
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_sim_r01k7.log | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_similarity2 -c 1 \
    -o gpurun_out/prof_sim_dbg7 -f python tools/bench_stats.py --only similarity --workloads C2 --repeats 1 > gpurun_out/ncu_sim_dbg7.log 2>&1
tail -1 gpurun_out/ncu_sim_dbg7.log | cut -c1-300
