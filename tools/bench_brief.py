#!/usr/bin/env python
"""Prints the essentials of a bench.py JSON line: value, ms/step, per-pass ms and GB/s."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c = d["config"]
print(f'{c["workload"]}: {d["value"]/1e6:.1f} M p-substeps/s  {d["ms_per_step"]:.3f} ms/step  e2e {(d["e2e"]["value"] or 0)/1e6:.1f} M/s  '
      f'pairs {c.get("pairs_searched")}/{c.get("pairs_kept")} unmirrored {c.get("pairs_unmirrored")}  launches {d["gpu_launches"]}  '
      f'substep roofline {d["roofline"]["substep"]["frac"]:.3f}  top {d["roofline"]["kernel"]} {d["roofline"]["frac"]:.3f}')
for k, v in d["passes"].items():
    print(f'   {k:16s} {v["ms_per_launch"]:8.4f} ms x{v["launches"]/d["steps"]:.0f}  share {v["share"]:.3f}  {v.get("gbs", "")}')
