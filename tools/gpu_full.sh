#!/bin/bash
# tests + bench + e2e breakdown
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py ) > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.2f e2e %.3e kernel_ms %.2f frac %.3f cpu %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value',0)))"
timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/e2e_probe.json'))
for s in d['e2e_breakdown']: print({k:(round(v,4) if not isinstance(v,dict) else {a:round(b,2) for a,b in v.items()}) for k,v in s.items()})
PY
