#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and a full ncu capture of the dominant kernels.
# Usage (here): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag] [workload]'
# Everything is written under gpurun_out/<tag>/ and merged back into the container.
set -u
TAG=${1:-r01}
WL=${2:-dam_break_1M}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1

echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.txt"

echo "== smoke"
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee "$OUT/smoke.txt"

echo "== bench ($WL)"
timeout 900 python bench.py --gpus 1 --steps 50 --warmup 3 --workload "$WL" > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -c 4000 "$OUT/bench.json"; tail -5 "$OUT/bench.err"

echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 --workload "$WL" > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
tail -c 1500 "$OUT/bench_ref.json"

echo "== ncu launch list (same command, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python bench.py --gpus 1 --steps 2 --warmup 3 --workload "$WL" --no-cpu-baseline --no-e2e > "$OUT/ncu_launches.log" 2>&1
tail -2 "$OUT/ncu_launches.log"

echo "== ncu --set full on the sweeps and the emit"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_density_lambda|k_apply_delta|k_green_stream|k_regroup|k_onesweep|k_reorder|k_begin_iteration' -s 45 -c 18 \
    -o "$OUT/prof" -f python bench.py --gpus 1 --steps 1 --warmup 3 --workload "$WL" --no-cpu-baseline --no-e2e > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log"
ls -la "$OUT"
