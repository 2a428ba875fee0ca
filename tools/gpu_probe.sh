#!/bin/bash
# one gpurun call: tests, bench, probes.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.log 2>&1
tail -4 gpurun_out/bench_n1.log
timeout 120 ./tools/pipe_probe > gpurun_out/pipe_probe.txt 2>&1
cat gpurun_out/pipe_probe.txt
timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err
cat gpurun_out/e2e_probe.json | head -80
tail -3 gpurun_out/e2e_probe.err
