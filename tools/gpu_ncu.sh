#!/bin/bash
# ncu --set full capture of selected kernels of one bench substep.  Usage: bash tools/gpu_ncu.sh <tag> <workload> <kernel regex> [skip] [count]
set -u
TAG=${1:-ncu}; WL=${2:-dam_break_1M}; RE=${3:-k_green_emit}; SKIP=${4:-0}; CNT=${5:-4}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s "$SKIP" -c "$CNT" \
    -o "$OUT/prof" -f python bench.py --gpus 1 --steps 1 --warmup 3 --workload "$WL" --no-cpu-baseline --no-e2e > "$OUT/ncu_full.log" 2>&1
tail -3 "$OUT/ncu_full.log"; ls -la "$OUT"
