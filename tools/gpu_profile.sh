#!/bin/bash
# One GPU-box visit for the profiles/ evidence: tools/gpu_profile.sh <tag> [workload]
#   1. launch list of two whole substeps (time + DRAM bytes per launch)        -> gpurun_out/<tag>/launches_<wl>.csv
#   2. ncu --set full, one substep's worth of the hot kernels, with source correlation -> gpurun_out/<tag>/prof_<wl>.ncu-rep
# (APBF_SIM_GRAPHS=0: the substeps are enqueued launch by launch, so that --launch-skip counts what the comments say)
set -u
TAG=${1:-r02}
WL=${2:-dam_break_1M}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export APBF_SIM_GRAPHS=0
B="python bench.py --gpus 1 --workload $WL --no-cpu-baseline --no-e2e --no-extra"
$B --steps 4 --warmup 3 > "$OUT/plain_$WL.json" 2> "$OUT/plain_$WL.err"
LPS=$(python -c "import json;d=json.loads(open('$OUT/plain_$WL.json').read().strip().splitlines()[-1]);print(d['gpu_launches']//d['steps'])")
echo "launches per substep: $LPS"
# upload + 3 warm-up substeps come first: skip them (set-up launches: ~6), then capture exactly two substeps
SKIP=$((3 * LPS + 8))
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $SKIP -c $((2 * LPS)) --csv \
    --log-file "$OUT/launches_$WL.csv" $B --steps 3 --warmup 3 > "$OUT/ncu_launches_$WL.log" 2>&1
tail -2 "$OUT/ncu_launches_$WL.log"
# the hot kernels of the SECOND substep (18 launches of these names per substep: 3 sort passes, reorder, emit, regroup, 4 x (prologue, T1, T2))
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_density_lambda|k_apply_delta|k_green_stream|k_regroup|k_onesweep|k_reorder|k_begin_iteration' -s 18 -c 18 \
    -o "$OUT/prof_$WL" -f $B --steps 1 --warmup 3 > "$OUT/ncu_full_$WL.log" 2>&1
tail -2 "$OUT/ncu_full_$WL.log"
ls -la "$OUT"
