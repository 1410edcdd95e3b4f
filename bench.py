#!/usr/bin/env python
"""bench.py -- particle-substeps/s of the APBF hot path (neighbour search + [adaptive kernel width] + 4-iteration PBF
solve) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                       (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...    the CPU arm: the oracle restatement of the reference shaders on the host cores

A step is one substep in the reference's order (source/pool.cpp:67-106) over the whole particle set.  `value` is timed
with the lists resident in HBM; `e2e` goes through the host-buffer entry points (upload -> substep -> download, the state
lives in pinned host memory between the steps).  Prints ONE JSON line on rank 0:

    value / ms_per_step / e2e / roofline / passes / clocks / gpu_launches      the workload BASELINE.json's metric is quoted on
    extra      (N = 1) the other BASELINE.json configs, the binary search and the operator-by-operator leg, a few steps each
    cpu_baseline, parity   (N = 1) the oracle on the host cores on a bounded sample, and the CUDA path checked against it
    extra / n_rank_parity  (N > 1) the 8 M-particles-per-GPU uniform scene in bricks, and N ranks == 1 rank bit for bit
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from apbf_b200 import scenes  # noqa: E402

METRIC = "particle-substeps/sec (search + 4-iter PBF solve)"
UNIT = "particle-substeps/s"


# one scene across N GPUs (weak scaling, 10^6 particles per GPU): the bricks are the halves of the search grid along z, then y,
# then x (apbf_b200/multi_gpu.py), so the dam-break grows along those axes and stays symmetric about the cutting planes
SLAB_DAM_BREAK = {2: dict(nx=100, ny=100, nz=200), 4: dict(nx=100, ny=200, nz=200, center_y=True),
                  8: dict(nx=100, ny=200, nz=200, blocks=2, center_y=True)}
# search grid of the timed dam breaks: equal extents on all axes, i.e. cubic cells of 4.8 r (scenes.dam_break, grid="cube")
DAM_BREAK_GRID = "cube"
# bounded sample of each workload for the CPU arm (the oracle needs ~10 us per particle-substep and thread)
CPU_SAMPLE = {"dam_break_1M": "dam_break_262k", "uniform_64": "uniform_64", "dam_break_1M_default_mode": "dam_break_64k_default_mode",
              "dam_break_1M_split_merge": "dam_break_64k_split_merge", "waterfall_16M": "waterfall_262k", "waterdrop_4M": "waterdrop_500k",
              "uniform_256": "uniform_64", "uniform_200": "uniform_64", "uniform_160": "uniform_64", "uniform_100": "uniform_64"}


def make_scene(name, world=1, res_log2=None):
    """BASELINE.json configs -> synthetic scenes (apbf_b200/scenes.py)"""
    rl = dict(res_log2=res_log2) if res_log2 else {}
    if name == "dam_break_1M" and world in SLAB_DAM_BREAK:
        return scenes.dam_break(adaptive=True, grid=DAM_BREAK_GRID, **{**SLAB_DAM_BREAK[world], **rl}), dict(adaptive=True, pairs_per_particle=150, slab=True)
    if name == "dam_break_1M":      # configs[1]: pool scene dam-break, 1M particles, adaptive kernel width
        return scenes.dam_break(100, 100, 100, adaptive=True, grid=DAM_BREAK_GRID, **rl), dict(adaptive=True, pairs_per_particle=150)
    if name == "dam_break_1M_default_mode":   # the reference's default adaptive mode: kernel width from the boundary distance
        # (pool.cpp:77-80) + update_transfers after the solver (pool.cpp:99-102, merge and split off); no spread_kernel_width
        return scenes.dam_break(100, 100, 100, adaptive=True, grid=DAM_BREAK_GRID, **rl), dict(adaptive=False, basic_pbf=False, update_transfers=True, pairs_per_particle=60)
    if name == "dam_break_1M_split_merge":   # the default mode with settings::merge / settings::split on (pool.cpp:73-75, :99-102;
        # SURVEY 8f row 3): particle_transfer after velocity_handling, merge / split decisions after the solver; room for 25 % copies
        return scenes.dam_break(100, 100, 100, adaptive=True, grid=DAM_BREAK_GRID, **rl), dict(adaptive=False, basic_pbf=False, update_transfers=True, transfers=True,
                                                                          pairs_per_particle=60, capacity_factor=1.25)
    if name == "dam_break_64k_split_merge":
        return scenes.dam_break(40, 40, 40, adaptive=True, grid=DAM_BREAK_GRID), dict(adaptive=False, basic_pbf=False, update_transfers=True, transfers=True,
                                                                 pairs_per_particle=60, capacity_factor=1.25)
    if name == "dam_break_64k_default_mode":
        return scenes.dam_break(40, 40, 40, adaptive=True, grid=DAM_BREAK_GRID), dict(adaptive=False, basic_pbf=False, update_transfers=True, pairs_per_particle=60)
    if name == "dam_break_64k":
        return scenes.dam_break(40, 40, 40, adaptive=True, grid=DAM_BREAK_GRID), dict(adaptive=True, pairs_per_particle=150)
    if name == "dam_break_262k":    # bounded sample of configs[1] for the CPU arm and the parity block: 64^3, the size of configs[0]
        return scenes.dam_break(64, 64, 64, adaptive=True, grid=DAM_BREAK_GRID), dict(adaptive=True, pairs_per_particle=150)
    if name == "uniform_64":        # configs[0]: uniform 64^3 block, fixed kernel width (jittered lattice)
        return scenes.uniform_block(64, jitter=0.1, shuffle=True, **rl), dict(adaptive=False, pairs_per_particle=40)
    if name == "uniform_32":
        return scenes.uniform_block(32, jitter=0.1, shuffle=True), dict(adaptive=False, pairs_per_particle=40)
    if name.startswith("uniform_"):  # configs[4]: uniform-block sweep, e.g. uniform_100 / 160 / 256 / 400
        return scenes.uniform_block(int(name.split("_")[1]), jitter=0.1, shuffle=True, **rl), dict(adaptive=False, pairs_per_particle=40, slab=True)
    if name == "waterfall_16M":     # configs[3]: 252^3 particles in the closed top pool, 11 collision boxes (waterfall.cpp:28-48)
        return scenes.waterfall(252, 252, 252, **rl), dict(adaptive=False, pairs_per_particle=40, slab=True)
    if name == "waterfall_262k":
        return scenes.waterfall(64, 64, 64), dict(adaptive=False, pairs_per_particle=40)
    if name == "waterfall_64k":
        return scenes.waterfall(40, 40, 40), dict(adaptive=False, pairs_per_particle=40)
    if name == "waterdrop_4M":      # configs[2]
        return scenes.waterdrop(204), dict(adaptive=True, pairs_per_particle=260)
    if name == "waterdrop_500k":
        return scenes.waterdrop(104), dict(adaptive=True, pairs_per_particle=330)
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(n, p_searched, p_kept, cells, bits, iters, adaptive):
    """compulsory DRAM bytes per launch of each pass (fused formulation), with the measured P."""
    passes = -(-bits // 8)
    per = {
        "hash_sort": 16 * n + 4 * n + 16 * n * passes,
        "reorder": 152 * n,
        "cell_ranges": 4 * n + 8 * cells,
        # adaptive: search and spread_kernel_width run fused (one pass of tests, only the kept pairs are written, first as
        # 8-byte stream entries and then as the grouped list); fixed widths: the same with kept == searched
        "emit_count": 16 * n + 4 * n + 8 * cells + 8 * p_kept,
        "emit_fill": 8 * p_kept + 4 * p_kept,     # stream entries in, 4-byte NB list out (the whole-scene path keeps no 8-byte list)
        "kw_spread": 8 * p_searched + 12 * n,
        "kw_compact": 8 * p_searched + 8 * p_kept,
        "box_collision": 36 * n,
        "density_lambda": 8 * p_kept + 56 * n,
        "apply_delta": 8 * p_kept + 40 * n,
        "update_transfers": 8 * p_kept + 48 * n,   # pairs + per particle: position, radius, old/new boundary distance, target radius, boundariness
    }
    substep = (196 + 16 * passes) * n + 16 * cells + 8 * p_searched + iters * (16 * p_kept + 132 * n)
    if adaptive:
        substep += 8 * p_searched + 8 * p_kept + 12 * n
    return per, substep


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if rows:
            sm = sorted(float(r[0]) for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][1])
            out["power_w_max"] = max(float(r[2]) for r in rows if r[2].replace(".", "").isdigit()) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].strip().lower() == "active" for r in rows)]
            out["samples"] = len(rows)
        return out


# ---- CPU arm: the oracle restatement of the reference's shaders on the host cores ---------------------------------------
def _oracle_setup(sample_name):
    from oracle import oracle as orc
    sc, meta = make_scene(sample_name)
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0 if meta["adaptive"] else 1
    s.mSmallestTargetRadius = sc.smallest_target_radius
    s.mMerge = s.mSplit = 1 if meta.get("transfers") else 0
    kw = dict(dims=sc.dims, basic_pbf=meta.get("basic_pbf", not meta["adaptive"]), solver_iterations=sc.solver_iterations, min_pos=sc.min_pos,
              max_pos=sc.max_pos, res_log2=sc.res_log2, box_min4=sc.box_min, box_max4=sc.box_max, cap=sc.n * meta["pairs_per_particle"], integrate=True,
              update_transfers=bool(meta.get("update_transfers")))
    if meta.get("transfers"):
        hidden_cap = int(sc.n * meta.get("capacity_factor", 1.0))
        kw.update(transfers=orc.Transfers(hidden_cap), hidden_cap=hidden_cap, split_duration=0.0)
    return orc, sc, meta, s, kw


def time_oracle(sample_name, steps, warmup, threads, budget_s=None):
    """particle-substeps/s of the oracle on `threads` host threads; with a budget the step count is cut so that the run ends in time"""
    orc, sc, meta, s, kw = _oracle_setup(sample_name)
    orc.set_threads(threads)
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    t0 = time.perf_counter()
    orc.substep(st, s, **kw)                 # the first step doubles as warm-up and as the probe for the budget
    t_step = time.perf_counter() - t0
    done_warm = 1
    if budget_s is not None:
        warmup = max(1, min(warmup, int(0.25 * budget_s / max(t_step, 1e-6))))
        steps = max(1, min(steps, int(0.75 * budget_s / max(t_step, 1e-6))))
    for _ in range(max(warmup - done_warm, 0)):
        orc.substep(st, s, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.substep(st, s, **kw)
    dt = time.perf_counter() - t0
    return sc.n * steps / dt, dt / steps, sc, steps, max(warmup, 1)


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself (Vulkan/GLSL, MSVC) cannot be
    built or run in this image, so this arm times the oracle restatement (kind = "port"), all host threads, on a bounded
    sample of the workload; steps and warm-up follow the command line unless that would take more than ~2.5 minutes."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = CPU_SAMPLE.get(args.workload, args.workload)
    value, sec_per_step, sc, steps, warm = time_oracle(sample, args.steps, args.warmup, threads, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32",
        "data": "synthetic", "config": {"workload": args.workload, "sample": sample, "particles": sc.n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample}: {sc.n} particles, {steps} substep(s) of the {args.workload} workload"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit_line(line)


def cpu_baseline_and_parity(gpu, workload, search):
    """rank 0, N = 1: the oracle on the host cores (all threads and one thread) on a bounded sample of the workload, and the
    CUDA path checked against it on that same sample (the oracle as the checker)"""
    from oracle import parity
    sample = CPU_SAMPLE.get(workload, workload)
    threads = os.cpu_count() or 1
    v_all, sec_all, ssc, steps_all, _ = time_oracle(sample, 4, 1, threads, budget_s=14.0)
    v_one, sec_one, _, steps_one, _ = time_oracle(sample, 1, 1, 1, budget_s=1.0)
    base = {"value": v_all, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample}: {ssc.n} particles, {steps_all} substeps of the {workload} workload on {threads} threads "
                      f"({sec_all * (steps_all + 1):.1f} s of CPU work) + {steps_one + 1} on one thread",
            "one_thread": {"value": v_one, "ms_per_step": sec_one * 1e3},
            "all_threads": {"value": v_all, "ms_per_step": sec_all * 1e3, "threads": threads, "speedup_over_one": v_all / v_one},
            "note": "searches, pair loops and per-particle passes of the port run on all threads (OpenMP); the 4-bit LSD sort and the "
                    "re-order gathers are serial, like one Vulkan queue's dispatch order"}
    sc, meta = make_scene(sample)
    ops = parity.operator_parity(gpu, sc, adaptive=meta["adaptive"], search=search, pairs_per_particle=meta["pairs_per_particle"], threads=threads)
    sub = parity.substep_parity(gpu, sc, adaptive=meta["adaptive"], substeps=2, search=search, pairs_per_particle=meta["pairs_per_particle"], threads=threads)
    sub0 = parity.substep_parity(gpu, sc, adaptive=meta["adaptive"], substeps=2, search=search, pairs_per_particle=meta["pairs_per_particle"], threads=threads,
                                 hk=0, gk=0)
    par = {"sample": f"{sample}: {sc.n} particles, same workload", "checker": "oracle/apbf_oracle.c (CPU restatement of the shaders; parity unpinned for the physics passes)",
           "bars": f"bit-exact: keys, order, cell tables, lists, pair list in order, kernel widths; accumulators {parity.ACC_UNITS} units of 2^-18 + 1e-5 rel; "
                   f"lambda 1e-5 rel where the accumulators agree; position shift of one iteration {parity.SHIFT_UNITS} units + {parity.SHIFT_REL:g} of the largest shift "
                   "(oracle/parity.py, BASELINE.md section 3)",
           "ok": bool(ops["ok"] and sub["ok"] and sub0["ok"]),
           "one_iteration_operator_by_operator": ops, "whole_substeps_gauss": sub, "whole_substeps_cubic": sub0}
    return base, par


# ---- GPU arm -----------------------------------------------------------------------------------------------------------
def _pinned_state(torch, gpu, arrays, capacity):
    """the scene's lists in pinned host memory, `capacity` rows each (what the e2e leg streams in and out every step)"""
    host = {}
    for name, dt, w in gpu.FIELDS:
        a = np.ascontiguousarray(arrays[name], dtype=dt).reshape(-1, w)
        t = torch.zeros((capacity, w), dtype=torch.int32 if dt != np.float32 else torch.float32).pin_memory()
        t[: a.shape[0]].copy_(torch.from_numpy(a.view(np.int32) if dt == np.uint32 else a))
        host[name] = t
    return host


def run_single(gpu, torch, workload, steps, warmup, *, device=0, search="green", res_log2=None, e2e_steps=0, clock_gpu=None, stats_pass=True):
    """one workload on one GPU through apbf_sim_*: device-resident timing, per-pass times, optional end-to-end leg"""
    sc, meta = make_scene(workload, 1, res_log2)
    ctx = gpu.Context(device=device, dims=sc.dims)
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius,
                     mMerge=1 if meta.get("transfers") else 0, mSplit=1 if meta.get("transfers") else 0)
    n, capacity = sc.n, int(sc.n * meta.get("capacity_factor", 1.0))
    sim = gpu.Sim(ctx, sc, capacity=capacity, neighbor_capacity=capacity * meta["pairs_per_particle"], integrate=True,
                  basic_pbf=meta.get("basic_pbf", not meta["adaptive"]), update_transfers=bool(meta.get("update_transfers")),
                  use_binary_search=(search == "binary"), transfers=bool(meta.get("transfers")))
    host = _pinned_state(torch, gpu, sc.arrays, capacity)
    sim.upload(host, n=n)
    sampler = ClockSampler(clock_gpu) if clock_gpu is not None else None
    for _ in range(warmup):
        sim.substep(1)
    torch.cuda.synchronize()
    # ---- the timed region: `steps` substeps, nothing else on the stream (the library replays them as captured CUDA graphs) ----
    launches0, replays0 = ctx.launch_count, sim.graph_replays()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.substep(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    replays = sim.graph_replays() - replays0
    # ---- the same substeps once more with a pair of CUDA events around every pass (72 event records per substep, which is why
    # they are kept out of the region above): per-pass durations for the roofline block ------------------------------------------
    prof_steps = max(3, min(steps, 20))
    ctx.profile(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(prof_steps):
        sim.substep(1)
    p1.record()
    torch.cuda.synchronize()
    ctx.profile(False)
    prof_ms = p0.elapsed_time(p1)
    prof = ctx.profile_read()
    stats = sim.stats()
    stats = dict(stats, list_state=ctx.list_state())
    clocks = sampler.stop() if sampler else None
    if meta["adaptive"] and stats_pass:
        # the fused search never builds the unpruned list; one more (untimed) substep counts what it would have held
        ctx.set_search_stats(True)
        sim.substep(1)
        stats = dict(stats, pairs_searched=sim.stats()["pairs_searched"])
        ctx.set_search_stats(False)
    flags = ctx.device_flags()

    # ---- end to end: the state lives in pinned host memory; every step copies all lists in, runs the substep, copies all lists
    # out again (the search re-orders every list, so the whole state comes back) ------------------------------------------------
    e2e = None
    if e2e_steps:
        n_now = sim.download(host)
        for _ in range(2):
            n_now = sim.step_host(host, n_now)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n_now = sim.step_host(host, n_now)        # all lists in, one substep, all lists out; synchronises
        e2e_s = time.perf_counter() - t0
        # the same three calls one after the other (no overlap of copies and work), for comparison
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.upload(host, n=n_now)
            sim.substep(1)
            n_now = sim.download(host)
        plain_s = time.perf_counter() - t1
        row = sum(w * 4 for _, _, w in gpu.FIELDS)
        e2e = {"value": n * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": row * n_now, "d2h_bytes_per_step": row * n_now,
               "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
               "what": "apbf_sim_step_host: all lists from pinned host memory in, one substep, all lists out again; the state evolves from step to "
                       "step; the small lists go up while positions are hashed and sorted, the lists the solver does not touch come down while it runs",
               "ms_per_step_upload_substep_download_in_turn": plain_s / e2e_steps * 1e3}
    sim.close()
    ctx.close()
    del host
    torch.cuda.empty_cache()
    return dict(sc=sc, meta=meta, n=n, ms=ms, steps=steps, launches=launches, prof=prof, prof_ms=prof_ms, prof_steps=prof_steps, graph_replays=replays,
                stats=stats, clocks=clocks, e2e=e2e, flags=flags)


def roofline_of(r, peak, workload, world=1):
    sc, meta, n, stats = r["sc"], r["meta"], r["n"], r["stats"]
    cells = 1 << (sc.res_log2 * sc.dims)
    bits = 4 * -(-(sc.res_log2 * sc.dims + 1) // 4)
    per_launch, substep_bytes = algorithmic_bytes(n, stats["pairs_searched"], stats["pairs_kept"], cells, bits, sc.solver_iterations, meta["adaptive"])
    ms_step = r["ms"] / r["steps"]
    timed = {k: v for k, v in r["prof"].items() if v[1] > 0}
    passes = {k: {"ms_per_launch": round(v[0] / v[1], 4), "launches": v[1], "share": round(v[0] / r.get("prof_ms", r["ms"]), 4),
                  **({"gbs": round(per_launch[k] / (v[0] / v[1] * 1e-3) / 1e9, 1), "frac": round(per_launch[k] / (v[0] / v[1] * 1e-3) / 1e9 / peak, 4)} if k in per_launch else {})}
              for k, v in timed.items()}
    top = max((k for k in timed if k in per_launch), key=lambda k: timed[k][0])
    top_ms = timed[top][0] / timed[top][1]
    achieved = per_launch[top] / (top_ms * 1e-3) / 1e9
    traffic, traffic_src, measured = None, None, None
    try:   # per-launch DRAM bytes of the passes from the committed ncu --set full capture of this workload (tools/ncu_traffic.py)
        tj = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{workload}.json")))
    except (OSError, ValueError):
        tj = None
    if tj and world == 1:
        tp = tj.get("passes", {})
        if top in tp:
            traffic, traffic_src = tp[top]["dram_bytes_per_launch"], tj["source"]
        if tj.get("substep_dram_bytes"):
            measured = {"dram_bytes": tj["substep_dram_bytes"], "achieved": tj["substep_dram_bytes"] / (ms_step * 1e-3) / 1e9,
                        "frac": tj["substep_dram_bytes"] / (ms_step * 1e-3) / 1e9 / peak,
                        "note": "DRAM bytes of one substep as ncu measured them (sum over the launch list) over this run's time"}
    roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": per_launch[top], "ms_per_launch": top_ms,
            "substep": {"algorithmic_bytes": substep_bytes, "achieved": substep_bytes / (ms_step * 1e-3) / 1e9,
                        "frac": substep_bytes / (ms_step * 1e-3) / 1e9 / peak,
                        "note": "SURVEY 8(d) byte formula (counts the unpruned pair list twice although the fused pass never moves it)"},
            "substep_measured_traffic": measured}
    return roof, passes


def run_operators(gpu, torch, workload, steps, search="green", device=0, fused_search=False):
    """one substep the way a scene of the reference calls it (pool.cpp:67-106): operator after operator through the drop-in
    surface, every list in its public format -- instead of the fused whole-scene call"""
    sc, meta = make_scene(workload)
    ctx = gpu.Context(device=device, dims=sc.dims)
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius)
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=sc.n * meta["pairs_per_particle"])
    vel = gpu.velocity_handling(ctx).set_data(L).set_acceleration((0.0, -10.0, 0.0))
    fused_search = fused_search and meta["adaptive"]    # search + spread_kernel_width as ONE operator (not a class of the reference)
    if search == "green":
        nbh = (gpu.neighborhood_green_spread if fused_search else gpu.neighborhood_green)(ctx).set_data(L).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2)
    else:
        nbh = (gpu.neighborhood_binary_search_spread if fused_search else gpu.neighborhood_binary_search)(ctx).set_data(L)
    nbh.set_range_scale(1.5 if meta["adaptive"] else 1.0)
    spread = gpu.spread_kernel_width(ctx).set_data(L)
    box = gpu.box_collision(ctx).set_data(L, sc.box_min, sc.box_max)
    inc = gpu.incompressibility(ctx).set_data(L)

    def substep():
        vel.apply(1.0 / 60.0)
        nbh.apply()
        if meta["adaptive"] and not fused_search:
            spread.apply()
        for _ in range(sc.solver_iterations):
            box.apply()
            inc.apply()

    for _ in range(3):
        substep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        substep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"ms_per_step": ms, "value": sc.n / (ms * 1e-3), "particles": sc.n, "pairs": L.pair_count(), "search": search,
           "what": ("velocity_handling, neighborhood_*_spread::apply (search + spread_kernel_width as one operator), " if fused_search else
                    "velocity_handling, neighborhood_*::apply, spread_kernel_width::apply, ") + "4 x (box_collision::apply, incompressibility::apply); every list in its public format"}
    del L
    ctx.close()
    torch.cuda.empty_cache()
    return out


def run_extras(gpu, torch, peak, device, quick):
    """the configs of BASELINE.json the headline is not quoted on, a few steps each (N = 1)"""
    out = {}
    plan = [("uniform_64", "green"), ("uniform_200", "green"), ("uniform_256", "green"), ("waterdrop_4M", "green"), ("waterfall_16M", "green"),
            ("dam_break_1M", "binary")]
    if quick:
        plan = [("uniform_64", "green")]
    for wl, search in plan:
        t0 = time.perf_counter()
        try:
            r = run_single(gpu, torch, wl, 6, 3, device=device, search=search, e2e_steps=0, stats_pass=False)
        except Exception as e:      # a config that does not fit is reported, not fatal
            out[wl + ("_binary_search" if search == "binary" else "")] = {"error": str(e)[:300]}
            continue
        roof, passes = roofline_of(r, peak, wl)
        ms = r["ms"] / r["steps"]
        out[wl + ("_binary_search" if search == "binary" else "")] = {
            "particles": r["n"], "pairs": r["stats"]["pairs_kept"], "res_log2": r["sc"].res_log2, "search": search, "ms_per_step": ms, "value": r["n"] / (ms * 1e-3),
            "substep_frac": roof["substep"]["frac"], "top_kernel": roof["kernel"], "top_kernel_frac": roof["frac"],
            "passes_ms": {k: v["ms_per_launch"] for k, v in passes.items()}, "gpu_launches": r["launches"], "device_flags": r["flags"],
            "wall_s": round(time.perf_counter() - t0, 1)}
    if not quick:
        for search in ("green", "binary"):
            out[f"dam_break_1M_operator_by_operator_{search}"] = run_operators(gpu, torch, "dam_break_1M", 6, search, device)
        out["dam_break_1M_operator_by_operator_green_fused_search"] = run_operators(gpu, torch, "dam_break_1M", 6, "green", device, fused_search=True)
    return out


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import apbf_b200 as gpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    if world > 1 and not args.replicas:
        import bench_multi
        return bench_multi.run(args, rank, world, local_rank, peak, peak_src, _emit_line, ClockSampler, make_scene, roofline_of, METRIC, UNIT)

    # ---- N = 1 (or N independent replicas) --------------------------------------------------------------------------------
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 20))
    r = run_single(gpu, torch, args.workload, args.steps, args.warmup, device=local_rank, search=args.search, res_log2=args.res_log2,
                   e2e_steps=e2e_steps, clock_gpu=local_rank if rank == 0 else None)
    ms, e2e = r["ms"], r["e2e"]
    if world > 1:
        t = torch.tensor([ms, -(e2e["value"] if e2e else 0.0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0].item())
        if e2e:
            e2e["value"] = -float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sc, meta, n, stats = r["sc"], r["meta"], r["n"], r["stats"]
    r["ms"] = ms
    roof, passes = roofline_of(r, peak, args.workload, world)
    roof["peak_source"] = peak_src
    if e2e:
        e2e["value"] *= world
    line = {
        "metric": METRIC, "value": n * world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": sc.name, "particles_per_gpu": n, "adaptive_kernel_width": meta["adaptive"],
                   "solver_iterations": sc.solver_iterations, "search": args.search, "res_log2": sc.res_log2,
                   "grid_cell": [round((h - l) / (1 << sc.res_log2), 3) for l, h in zip(sc.min_pos, sc.max_pos)],
                   "pairs_searched": stats["pairs_searched"], "pairs_kept": stats["pairs_kept"],
                   "pairs_unmirrored": stats["pairs_unmirrored"], "device_flags": r["flags"], "list_state": stats.get("list_state"),
                   "multi_gpu": "single" if world == 1 else "independent replicas",
                   "l2": "working set (lists + pair list) exceeds the 126 MB L2" if (76 * n + 8 * stats["pairs_searched"]) > 126e6
                         else "working set fits L2; no flush between steps"},
        "gpu_launches": r["launches"],
        "launch_mode": {"substeps_replayed_as_cuda_graph": r["graph_replays"], "of": args.steps,
                        "passes_timed_on": f"{r['prof_steps']} further substeps with CUDA events around every pass ({r['prof_ms'] / r['prof_steps']:.3f} ms/substep with the events)"},
        "e2e": e2e if e2e else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 0},
        "roofline": roof, "passes": passes, "clocks": r["clocks"],
    }
    if world == 1 and not args.no_extra:
        line["extra"] = run_extras(gpu, torch, peak, local_rank, args.quick_extra)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity(gpu, args.workload, args.search)
    _emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def _emit_line(line):
    """the ONE JSON line goes to the process's original stdout (see main(): fd 1 is pointed at stderr while the bench runs,
    because libraries print there -- NCCL's version banner for one)"""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    os.write(fd, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dam_break_1M")
    ap.add_argument("--res-log2", type=int, default=None, help="override the scene's search-grid resolution (any value gives the same lists)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (oracle timing + parity block)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the sweep over the other BASELINE.json configs")
    ap.add_argument("--quick-extra", action="store_true", help="the sweep with its smallest config only")
    ap.add_argument("--search", default="green", choices=["green", "binary"],
                    help="neighborhood_green (default, pool.cpp NEIGHBORHOOD_TYPE 1) or neighborhood_binary_search (type 3; 1 GPU only)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent copies of the 1-GPU scene instead of one scene in bricks")
    ap.add_argument("--mg-python", action="store_true", help="N > 1: drive the slab protocol from Python (apbf_b200/multi_gpu.py) instead of the library's own loop")
    ap.add_argument("--no-mg-extra", action="store_true", help="N > 1: skip the 8 M-particles-per-GPU run and the N-rank parity check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
