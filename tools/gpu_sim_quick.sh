#!/bin/bash
# similarity kernel: parity subset + throughput (C2, C3)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "simil or smoke or golden or dropin or pytrimal" 2>&1 | tail -2
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_sim_quick.log | cut -c1-330
