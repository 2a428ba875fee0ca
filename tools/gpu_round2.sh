#!/bin/bash
# GPU tests (all), consumer timings, trim() wall time through pytrimal, bench both arms
TAG=${1:-r01d}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python tools/bench_stats.py --only consumers --workloads C3,C4 --repeats 3 > gpurun_out/stats_cons_$TAG.log 2>&1; cut -c1-1800 gpurun_out/stats_cons_$TAG.log
timeout 900 python tools/trim_wall.py --configs C2,C4,C5 > gpurun_out/trim_wall_$TAG.log 2>&1; cat gpurun_out/trim_wall_$TAG.log | cut -c1-700
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-1500
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_ref.log 2>&1; tail -1 gpurun_out/bench_${TAG}_ref.log | cut -c1-500
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
