#!/bin/bash
# ncu --set full of K1 at full C4 size and of K4 at C2 size + the bench launch list
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_identity2 -s 3 -c 1 \
    -o gpurun_out/prof_id2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_id2_$TAG.log 2>&1
tail -2 gpurun_out/ncu_id2_$TAG.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_similarity2 -c 1 \
    -o gpurun_out/prof_sim_$TAG -f python tools/bench_stats.py --only similarity --workloads C2 --repeats 1 > gpurun_out/ncu_sim_$TAG.log 2>&1
tail -2 gpurun_out/ncu_sim_$TAG.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_$TAG.log 2>&1
tail -1 gpurun_out/launches_$TAG.log | cut -c1-200
ls -la gpurun_out/ | tail -8
