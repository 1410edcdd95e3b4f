#!/usr/bin/env python
"""One substep the way a scene of the reference calls it (pool.cpp:67-106): operator after operator through the drop-in
surface -- velocity_handling, neighborhood_green | neighborhood_binary_search, spread_kernel_width, 4 x (box_collision,
incompressibility) -- instead of the fused whole-scene call that bench.py times.  Prints ms per substep and per operator.
Usage (GPU box): python tools/bench_operators.py [workload] [green|binary] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import apbf_b200 as gpu  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "dam_break_1M"
search = sys.argv[2] if len(sys.argv) > 2 else "green"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
sc, meta = bench.make_scene(wl)
ctx = gpu.Context(dims=sc.dims)
ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius)
L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=sc.n * meta["pairs_per_particle"])
vel = gpu.velocity_handling(ctx).set_data(L).set_acceleration((0.0, -10.0, 0.0))
if search == "green":
    nbh = gpu.neighborhood_green(ctx).set_data(L).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2)
else:
    nbh = gpu.neighborhood_binary_search(ctx).set_data(L)
nbh.set_range_scale(1.5 if meta["adaptive"] else 1.0)
spread = gpu.spread_kernel_width(ctx).set_data(L)
box = gpu.box_collision(ctx).set_data(L, sc.box_min, sc.box_max)
inc = gpu.incompressibility(ctx).set_data(L)


def substep():
    vel.apply(1.0 / 60.0)
    nbh.apply()
    if meta["adaptive"]:
        spread.apply()
    for _ in range(sc.solver_iterations):
        box.apply()
        inc.apply()


for _ in range(3):
    substep()
torch.cuda.synchronize()
ctx.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    substep()
e1.record()
torch.cuda.synchronize()
ctx.profile(False)
ms = e0.elapsed_time(e1) / steps
print(f"{wl} [{search}] operator by operator: {ms:.3f} ms/substep = {sc.n / ms / 1e3:.1f} M particle-substeps/s, pairs {L.pair_count()}")
for k, (t, c) in ctx.profile_read().items():
    if c:
        print(f"   {k:18s} {t / steps:8.4f} ms/substep  ({c // steps} launches)")
