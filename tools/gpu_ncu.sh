#!/bin/bash
# ncu --set full of the identity kernel (bench.py --rows R), report back as gpurun_out/prof_<tag>.ncu-rep
TAG=${1:-v2}; ROWS=${2:-16384}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_identity -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f python bench.py --rows $ROWS --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/
