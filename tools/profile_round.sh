#!/bin/bash
# tools/profile_round.sh TAG -- on the GPU box (one GPU): the ncu evidence of a round.
#   1. launch list of a short bench.py run (per-launch gpu__time_duration, --clock-control none)
#   2. ncu --set full of one K1 launch in threshold mode (tcu_representatives, C4) and one in float mode;
#      one launch each of K0, the mirror / relayout pass and the clustering kernel
#   3. launch list of the spurious / gaps kernels at C5
TAG=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c4.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-similarity > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
cat > /tmp/reps.py <<'PY'
import sys; sys.path.insert(0, ".")
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import CONFIGS, synthetic_msa
n, L, seed = CONFIGS["C4"]
m = synthetic_msa(n, L, seed)
with pb.DeviceAlignment(m) as d:
    for _ in range(2):
        print(len(d.representatives(0.8, indet=ord("X"))))
    d.identity_on_device(ord("X"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_identity2 -s 1 -c 2 \
    -o gpurun_out/${TAG}_ncu_identity2_c4 -f python /tmp/reps.py > gpurun_out/${TAG}_ncu_identity2.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_identity2_c4.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_identity2_c4_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_ncu_identity2_c4_raw.csv > gpurun_out/${TAG}_ncu_identity2_c4.txt 2>&1
#   2b. one launch each of the pack, mirror / relayout and clustering kernels of the same call
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pack_planes|k_bits_rows|k_greedy_clusters' -c 3 \
    -o gpurun_out/${TAG}_ncu_consumers_c4 -f python /tmp/reps.py > gpurun_out/${TAG}_ncu_consumers.log 2>&1
ncu -i gpurun_out/${TAG}_ncu_consumers_c4.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_consumers_c4_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_ncu_consumers_c4_raw.csv > gpurun_out/${TAG}_ncu_consumers_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_stats_c5.csv python tools/bench_stats.py --only gaps,spurious --workloads C5 --repeats 1 > /dev/null 2>&1
ls -la gpurun_out | tail -8
