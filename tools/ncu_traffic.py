#!/usr/bin/env python
"""DRAM traffic of the passes from ncu captures -> profiles/traffic_<workload>.json (read by bench.py: roofline.traffic and
roofline.substep_measured_traffic).

    python tools/ncu_traffic.py <launches.csv> <workload> <tag> <substeps captured>

<launches.csv>: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv` over
whole substeps of `bench.py --workload <workload>` (tools/gpu_profile.sh).  Launches are keyed by their FULL kernel name, so
that template variants that return at once (k_apply_delta<1, true> on a list without unmirrored pairs, the overflow fall-back
of the emit) do not dilute the averages of the ones that do the work."""
import csv
import json
import os
import sys
from collections import defaultdict

PASS_OF = {"k_green_stream": "emit_count", "k_regroup": "emit_fill", "k_regroup_tma": "emit_fill", "k_density_lambda": "density_lambda",
           "k_apply_delta": "apply_delta", "k_begin_iteration": "box_collision", "k_reorder": "reorder", "k_onesweep": "hash_sort"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}

path, workload, tag, substeps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = [r for r in csv.reader(open(path, errors="ignore")) if r and r[0].strip('"').isdigit()]
# long format: ID, PID, Process, Host, Kernel Name, Context, Stream, Block, Grid, CC, Section, Metric Name, Unit, Value
per_launch = defaultdict(dict)
names = {}
for r in rows:
    lid, name, metric, unit, val = r[0], r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    names[lid] = name
    per_launch[lid][metric] = val * UNIT.get(unit, 1.0)
by_name = defaultdict(list)
total_bytes, total_us = 0.0, 0.0
for lid, m in per_launch.items():
    b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    t = m.get("gpu__time_duration.sum", 0.0)
    by_name[names[lid]].append((b, t))
    total_bytes += b
    total_us += t
out = {"workload": workload, "source": f"profiles/{tag}_launches_{workload}.csv (ncu --clock-control none, every launch of {substeps} substeps; per launch, keyed by the full kernel name)",
       "substeps_captured": substeps, "substep_dram_bytes": total_bytes / substeps, "substep_kernel_time_us_serialised": total_us / substeps,
       "passes": {}, "kernels": {}}
for name, v in sorted(by_name.items(), key=lambda kv: -sum(t for _, t in kv[1])):
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    out["kernels"][short] = {"launches": len(v), "dram_bytes_per_launch": sum(b for b, _ in v) / len(v), "us_per_launch": sum(t for _, t in v) / len(v),
                             "share_of_kernel_time": sum(t for _, t in v) / total_us}
    base = short.split("<")[0]
    if base in PASS_OF:
        p = PASS_OF[base]
        cur = out["passes"].get(p)
        cand = {"kernel": short, "dram_bytes_per_launch": sum(b for b, _ in v) / len(v), "us_per_launch": sum(t for _, t in v) / len(v), "launches_captured": len(v)}
        if cur is None or cand["us_per_launch"] > cur["us_per_launch"]:   # the variant that does the work
            out["passes"][p] = cand
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"traffic_{workload}.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: out[k] for k in ("substep_dram_bytes", "substep_kernel_time_us_serialised")}), json.dumps(out["passes"], indent=1))
