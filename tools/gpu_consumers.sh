#!/bin/bash
# identity-matrix consumers: timings, ncu launch list, ncu --set full of K5..K8 + histogram
TAG=${1:-cons}
mkdir -p gpurun_out
timeout 600 python tools/bench_stats.py --only consumers --workloads C4 --repeats 3 | tee gpurun_out/stats_$TAG.log | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/bench_stats.py --only consumers --workloads C4 --repeats 1 > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_identity_bits|k_row_stats|k_mis_scan|k_mis_resolve|k_row_lengths' -c 12 \
    -o gpurun_out/prof_$TAG -f python tools/bench_stats.py --only consumers --workloads C4 --repeats 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
ls -la gpurun_out
