#!/bin/bash
# all GPU tests + smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 ) > gpurun_out/pytest_gpu_r01k.log 2>&1; tail -5 gpurun_out/pytest_gpu_r01k.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/trim_wall.py --configs C3 2>/dev/null | grep '^{' | cut -c1-400
