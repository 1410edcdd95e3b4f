#!/bin/bash
# launch list (time + DRAM bytes per launch) of two substeps: tools/gpu_launches.sh <tag> [workload]  -> gpurun_out/<tag>_launches_<wl>.csv
set -u
TAG=${1:-r02}; WL=${2:-dam_break_1M}
export APBF_SIM_GRAPHS=0
B="python bench.py --gpus 1 --workload $WL --no-cpu-baseline --no-e2e --no-extra"
$B --steps 4 --warmup 3 > "gpurun_out/${TAG}_plain_$WL.json" 2> "gpurun_out/${TAG}_plain_$WL.err"
LPS=$(python -c "import json;d=json.loads(open('gpurun_out/${TAG}_plain_$WL.json').read().strip().splitlines()[-1]);print(d['gpu_launches']//d['steps'])")
echo "launches per substep: $LPS"
SKIP=$((3 * LPS + 8))
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c $((2 * LPS)) --csv \
    --log-file "gpurun_out/${TAG}_launches_$WL.csv" $B --steps 3 --warmup 3 > "gpurun_out/${TAG}_ncu_launches_$WL.log" 2>&1
tail -2 "gpurun_out/${TAG}_ncu_launches_$WL.log"
