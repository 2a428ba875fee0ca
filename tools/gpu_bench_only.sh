#!/bin/bash
# default bench.py run (N=1)
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/bench_r01k_n1.log 2>&1; tail -5 gpurun_out/bench_r01k_n1.log | cut -c1-400
