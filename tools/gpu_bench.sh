#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/id2_debug.py > gpurun_out/id2_debug.log 2>&1; tail -2 gpurun_out/id2_debug.log
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_v2.log 2>&1; tail -1 gpurun_out/bench_v2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.2f e2e %.3e kernel_ms %.2f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']))"
( timeout 600 python bench.py --no-cpu-baseline --rows 16384 ) > gpurun_out/bench_v2_16k.log 2>&1; tail -1 gpurun_out/bench_v2_16k.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('16k: value %.3e kernel_ms %.2f frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
