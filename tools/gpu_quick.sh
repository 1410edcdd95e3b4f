set -u
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
for wl in dam_break_1M uniform_64 ${2:-}; do
timeout 600 python bench.py --steps 30 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2>$OUT/bench_$wl.err; python tools/bench_brief.py $OUT/bench_$wl.json; tail -3 $OUT/bench_$wl.err
done
