#!/bin/bash
# torchrun wrapper: rank 0 runs under ncu (launch list: duration per launch of the slab kernels), the other ranks run plain.
# Usage: python -m torch.distributed.run --no-python ... bash tools/mg_ncu_rank0.sh <out.csv> <bench args...>
OUT=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
    exec ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mgl|k_density_lambda|k_apply_delta|k_begin_iteration' -s 400 -c 400 --csv --log-file "$OUT" python bench.py "$@"
else
    exec python bench.py "$@"
fi
