#!/bin/bash
# bench variants in one GPU visit: tools/gpu_variants.sh <tag>; results under gpurun_out/<tag>_*.json
TAG=${1:-v}
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e"
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3),'ms', {k:v['ms_per_launch'] for k,v in d['passes'].items() if v['share']>0.02})
except Exception as e: print('ERR',e)
")"; }
run base $B
APBF_REGROUP_TMA=1 run tma $B
run res8 $B --res-log2 8
run res6 $B --res-log2 6
run u256 $B --workload uniform_256 --steps 6
run u256_res8 $B --workload uniform_256 --steps 6 --res-log2 8
APBF_REGROUP_TMA=1 run u256_tma $B --workload uniform_256 --steps 6
run u64 $B --workload uniform_64
run u64_res6 $B --workload uniform_64 --res-log2 6
run wd4m $B --workload waterdrop_4M --steps 6
run wd4m_res7 $B --workload waterdrop_4M --steps 6 --res-log2 7
