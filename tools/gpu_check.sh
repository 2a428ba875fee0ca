#!/bin/bash
# full GPU test suite + smoke + similarity throughput (C2, C3)
TAG=${1:-chk}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_sim_$TAG.log | cut -c1-600
