#!/bin/bash
# A/B of env-var controlled variants: bench kernel_ms at C4 (and 16k rows with FULL=1)
mkdir -p gpurun_out
run() { # label, env...
  local label="$1"
  for rows in 0 ${FULL:+16384}; do
    env $1 timeout 600 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 5 --rows $rows 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$label rows=$rows: value %.3e kernel_ms %.3f frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
}
timeout 300 python tools/id2_debug.py 2>&1 | tail -1
for v in "$@"; do run "$v"; done
