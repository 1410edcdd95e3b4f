#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
Usage: python tools/ncu_lines.py <rep> <kernel regex> [launch index] [top N]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ie, ss, ns = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Warp Stall Sampling (Not-issued Samples)")
inst, stall, text = defaultdict(int), defaultdict(int), {}
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0].strip():
        try:
            cur = int(r[0]); text[cur] = r[1]
        except ValueError:
            cur = None  # a new file/function header
    if cur is None or not r[2].strip():
        continue
    try:
        inst[cur] += int(r[ie]); stall[cur] += int(r[ss])
    except ValueError:
        pass
ti, ts = sum(inst.values()) or 1, sum(stall.values()) or 1
print(f"{rows[1][1][:90] if len(rows) > 1 and len(rows[1]) > 1 else pat}: {ti} warp instructions, {ts} stall samples")
for ln in sorted(inst, key=lambda k: -(inst[k] / ti + stall[k] / ts))[:top]:
    print(f"{inst[ln] / ti * 100:5.1f}% inst {stall[ln] / ts * 100:5.1f}% samples  L{ln}: {text.get(ln, '').strip()[:120]}")
