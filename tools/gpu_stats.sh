#!/bin/bash
# GPU tests + default bench + the other statistics' throughput (tools/bench_stats.py)
TAG=${1:-stats}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-600
( timeout 900 python tools/bench_stats.py --repeats ${2:-2} ) > gpurun_out/stats_$TAG.log 2>&1; cat gpurun_out/stats_$TAG.log
