#!/bin/bash
# all GPU tests, smoke, similarity throughput, bench (N=1)
TAG=${1:-r01k}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | tee gpurun_out/stats_sim_$TAG.log | cut -c1-600
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-2500
