#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,temperature.gpu --format=csv,noheader
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_sim_r01k2.log | cut -c1-330
for i in 1 2; do
timeout 600 python -m pytest tests -m gpu -x -q --capture=sys --timeout=150 > gpurun_out/dbg_pytest_all_$i.log 2>&1
grep -v "^  File" gpurun_out/dbg_pytest_all_$i.log | grep -i "passed\|failed\|abort\|terminate\|free()\|malloc\|corrupt\|what" | head -5 | cut -c1-300
done
