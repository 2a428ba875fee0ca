#!/bin/bash
# tools/profile_sim.sh TAG -- ncu --set full with source counters of one k_similarity2 launch at C2
TAG=${1:-sim}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_similarity2 -s 1 -c 1 \
    -o gpurun_out/${TAG}_ncu_similarity2_c2 -f python tools/bench_stats.py --only similarity --workloads C2 --repeats 2 > gpurun_out/${TAG}_ncu_sim.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_similarity2_c2.ncu-rep --page raw --csv > gpurun_out/${TAG}_sim_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_ncu_similarity2_c2.ncu-rep --page source --csv > gpurun_out/${TAG}_sim_source.csv 2>/dev/null
python tools/ncu_source.py gpurun_out/${TAG}_sim_source.csv 30 > gpurun_out/${TAG}_sim_source.txt 2>&1
rm -f gpurun_out/${TAG}_ncu_similarity2_c2.ncu-rep
timeout 200 python tools/bench_stats.py --only similarity --workloads C2,C3 --repeats 2 | cut -c1-330
