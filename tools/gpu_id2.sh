#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/id2_debug.py > gpurun_out/id2_debug.log 2>&1; echo "debug rc=$?"
cat gpurun_out/id2_debug.log | tail -30
if grep -q "^OK" gpurun_out/id2_debug.log; then
  ( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
  ( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_v2.log 2>&1; tail -2 gpurun_out/bench_v2.log
  ( TCU_IDENTITY_IMPL=v1 timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_v1.log 2>&1; tail -2 gpurun_out/bench_v1.log
fi
