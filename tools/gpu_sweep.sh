#!/bin/bash
# One GPU-box visit over every BASELINE.json config that fits one GPU (bench lines only, no profiler):
#   gpurun --timeout 1500 -- 'bash tools/gpu_sweep.sh [tag]'
# Writes gpurun_out/<tag>/bench_<workload>.json; tools/bench_brief.py prints one summary line per file.
set -u
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for wl in dam_break_1M dam_break_1M_default_mode dam_break_1M_split_merge uniform_64 uniform_160 uniform_256 waterdrop_4M waterfall_16M; do
    timeout 400 python bench.py --steps 30 --warmup 3 --workload "$wl" --no-cpu-baseline > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err" \
        && python tools/bench_brief.py "$OUT/bench_$wl.json" | head -1 || tail -3 "$OUT/bench_$wl.err"
done
