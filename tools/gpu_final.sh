#!/bin/bash
# round-end evidence: all GPU tests, smoke, the statistics' throughput, bench (both arms),
# trim() wall time through pytrimal, ncu captures of K1 (C4) and K4 (C2), bench launch list
TAG=${1:-r01k}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout=150 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/bench_stats.py --only similarity,gaps,spurious --workloads C2,C3,C5 --repeats 2 > gpurun_out/stats_$TAG.log 2>&1; cut -c1-260 gpurun_out/stats_$TAG.log
( timeout 600 python bench.py ) > gpurun_out/bench_${TAG}_n1.log 2>&1; tail -1 gpurun_out/bench_${TAG}_n1.log | cut -c1-300
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_${TAG}_ref.log 2>&1; tail -1 gpurun_out/bench_${TAG}_ref.log | cut -c1-300
timeout 600 python tools/trim_wall.py --configs C2,C3,C4,C5 > gpurun_out/trim_wall_$TAG.log 2>&1; cut -c1-420 gpurun_out/trim_wall_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_identity2 -s 3 -c 1 \
    -o gpurun_out/prof_id2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_id2_$TAG.log 2>&1
tail -1 gpurun_out/ncu_id2_$TAG.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_similarity2 -c 1 \
    -o gpurun_out/prof_sim_$TAG -f python tools/bench_stats.py --only similarity --workloads C2 --repeats 1 > gpurun_out/ncu_sim_$TAG.log 2>&1
tail -1 gpurun_out/ncu_sim_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_$TAG.log 2>&1
tail -1 gpurun_out/launches_$TAG.log | cut -c1-200
