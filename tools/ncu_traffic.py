#!/usr/bin/env python
"""DRAM traffic per launch of every profiled pass from an `ncu --set full` report -> profiles/traffic_<workload>.json.
bench.py puts the entry of its dominant pass into roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum, mean
over the captured launches).  Usage: python tools/ncu_traffic.py <prof.ncu-rep> <workload> <tag>"""
import csv
import json
import os
import subprocess
import sys

PASS_OF = {"k_green_stream": "emit_count", "k_regroup": "emit_fill", "k_density_lambda": "density_lambda", "k_apply_delta": "apply_delta",
           "k_begin_iteration": "box_collision", "k_reorder": "reorder", "k_onesweep": "hash_sort"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rep, workload, tag = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
acc = {}
for r in rows[2:]:
    for k, p in PASS_OF.items():
        if k + "<" in r[ik] or k + "(" in r[ik]:
            b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
            acc.setdefault(p, {"kernel": k, "bytes": [], "note": "k_onesweep: one of the radix passes only" if k == "k_onesweep" else None})["bytes"].append(b)
out = {"workload": workload, "source": f"profiles/{tag}_ncu_full_{workload}.csv (ncu --set full --clock-control none, per launch)", "passes": {}}
for p, v in acc.items():
    out["passes"][p] = {"kernel": v["kernel"], "dram_bytes_per_launch": sum(v["bytes"]) / len(v["bytes"]), "launches_captured": len(v["bytes"])}
    if v["note"]:
        out["passes"][p]["note"] = v["note"]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"traffic_{workload}.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out["passes"], indent=1))
