#!/usr/bin/env python
"""bench.py -- particle-substeps/s of the APBF hot path (neighbour search + [adaptive kernel width] + 4-iteration PBF
solve) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                       (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...    the CPU arm: the oracle restatement of the reference shaders on the host cores

A step is one substep in the reference's order (source/pool.cpp:67-106) over the whole particle set.  `value` is timed
with the lists resident in HBM; `e2e` goes through the host-buffer entry points (upload -> substep -> download).
Prints ONE JSON line on rank 0.
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
SLAB_DAM_BREAK = {2: dict(nx=100, ny=100, nz=200), 4: dict(nx=100, ny=200, nz=200, center_y=True, res_log2=8),
                  8: dict(nx=100, ny=200, nz=200, blocks=2, center_y=True)}


def make_scene(name, world=1):
    """BASELINE.json configs -> synthetic scenes (apbf_b200/scenes.py)"""
    if name == "dam_break_1M" and world in SLAB_DAM_BREAK:
        return scenes.dam_break(adaptive=True, **SLAB_DAM_BREAK[world]), dict(adaptive=True, pairs_per_particle=150, slab=True)
    if name == "dam_break_1M":      # configs[1]: pool scene dam-break, 1M particles, adaptive kernel width
        return scenes.dam_break(100, 100, 100, adaptive=True), dict(adaptive=True, pairs_per_particle=150)
    if name == "dam_break_1M_default_mode":   # the reference's default adaptive mode: kernel width from the boundary distance
        # (pool.cpp:77-80) + update_transfers after the solver (pool.cpp:99-102, merge and split off); no spread_kernel_width
        return scenes.dam_break(100, 100, 100, adaptive=True), dict(adaptive=False, basic_pbf=False, update_transfers=True, pairs_per_particle=60)
    if name == "dam_break_1M_split_merge":   # the default mode with settings::merge / settings::split on (pool.cpp:73-75, :99-102;
        # SURVEY 8f row 3): particle_transfer after velocity_handling, merge / split decisions after the solver; room for 25 % copies
        return scenes.dam_break(100, 100, 100, adaptive=True), dict(adaptive=False, basic_pbf=False, update_transfers=True, transfers=True,
                                                                    pairs_per_particle=60, capacity_factor=1.25)
    if name == "dam_break_64k_split_merge":
        return scenes.dam_break(40, 40, 40, adaptive=True), dict(adaptive=False, basic_pbf=False, update_transfers=True, transfers=True,
                                                                 pairs_per_particle=60, capacity_factor=1.25)
    if name == "dam_break_64k_default_mode":
        return scenes.dam_break(40, 40, 40, adaptive=True), dict(adaptive=False, basic_pbf=False, update_transfers=True, pairs_per_particle=60)
    if name == "dam_break_64k":     # bounded sample of the same workload for the CPU arm
        return scenes.dam_break(40, 40, 40, adaptive=True), dict(adaptive=True, pairs_per_particle=150)
    if name == "uniform_64":        # configs[0]: uniform 64^3 block, fixed kernel width (jittered lattice)
        return scenes.uniform_block(64, jitter=0.1, shuffle=True), dict(adaptive=False, pairs_per_particle=40)
    if name == "uniform_32":
        return scenes.uniform_block(32, jitter=0.1, shuffle=True), dict(adaptive=False, pairs_per_particle=40)
    if name.startswith("uniform_"):  # configs[4]: uniform-block sweep, e.g. uniform_100 / 160 / 256
        return scenes.uniform_block(int(name.split("_")[1]), jitter=0.1, shuffle=True), dict(adaptive=False, pairs_per_particle=40)
    if name == "waterfall_16M":     # configs[3]: 252^3 particles in the closed top pool, 11 collision boxes (waterfall.cpp:28-48)
        return scenes.waterfall(252, 252, 252), dict(adaptive=False, pairs_per_particle=40)
    if name == "waterfall_64k":
        return scenes.waterfall(40, 40, 40), dict(adaptive=False, pairs_per_particle=40)
    if name == "waterdrop_4M":      # configs[2]
        return scenes.waterdrop(204), dict(adaptive=True, pairs_per_particle=260)
    raise SystemExit(f"unknown workload {name}")


def settings_for(mod, adaptive):
    s = mod.default_settings() if hasattr(mod, "default_settings") else None
    return s


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


# ---- CPU arm: the oracle restatement of the reference's shaders, all host threads ------------------------------------
def time_oracle(sample_name, steps, warmup, threads):
    from oracle import oracle as orc
    sc, meta = make_scene(sample_name)
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0 if meta["adaptive"] else 1
    s.mSmallestTargetRadius = sc.smallest_target_radius
    s.mMerge = s.mSplit = 1 if meta.get("transfers") else 0
    orc.set_threads(threads)
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    cap = sc.n * meta["pairs_per_particle"]
    kw = dict(dims=sc.dims, basic_pbf=meta.get("basic_pbf", not meta["adaptive"]), solver_iterations=sc.solver_iterations, min_pos=sc.min_pos,
              max_pos=sc.max_pos, res_log2=sc.res_log2, box_min4=sc.box_min, box_max4=sc.box_max, cap=cap, integrate=True,
              update_transfers=bool(meta.get("update_transfers")))
    if meta.get("transfers"):
        hidden_cap = int(sc.n * meta.get("capacity_factor", 1.0))
        kw.update(transfers=orc.Transfers(hidden_cap), hidden_cap=hidden_cap, split_duration=0.0)
    for _ in range(warmup):
        orc.substep(st, s, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.substep(st, s, **kw)
    dt = time.perf_counter() - t0
    return sc.n * steps / dt, dt / steps, sc


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself (Vulkan/GLSL, MSVC) cannot be
    built or run in this image, so this arm times the oracle restatement (kind = "port") on a bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = {"dam_break_1M": "dam_break_64k", "uniform_64": "uniform_32",
              "dam_break_1M_default_mode": "dam_break_64k_default_mode",
              "dam_break_1M_split_merge": "dam_break_64k_split_merge"}.get(args.workload, args.workload)
    steps = max(1, min(args.steps, 150))     # ~0.4 s per step of the 64k-particle sample on 16 threads
    warm = min(args.warmup, 3)
    value, sec_per_step, sc = time_oracle(sample, steps, warm, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32",
        "data": "synthetic", "config": {"workload": args.workload, "sample": sample, "particles": sc.n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample}: {sc.n} particles, {steps} substep(s) of the {args.workload} workload"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit_line(line)


# ---- GPU arm -----------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import apbf_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sc, meta = make_scene(args.workload, 1 if args.replicas else world)
    slab = world > 1 and bool(meta.get("slab"))
    ctx = apbf_b200.Context(device=local_rank, dims=sc.dims)
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius,
                     mMerge=1 if meta.get("transfers") else 0, mSplit=1 if meta.get("transfers") else 0)
    if slab:
        # this rank's brick of the scene: the particles whose cell key starts with the rank's bits
        from apbf_b200 import multi_gpu
        owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
        arrays = {k: np.ascontiguousarray(v[owner == rank]) for k, v in sc.arrays.items()}
        n = len(arrays["position"])
        arrays["index_list"] = np.arange(n, dtype=np.uint32)
        n_total = sc.n
        del owner
        ghost_cap = 400_000
        capacity = int(n * 1.25) + ghost_cap
    else:
        arrays, n, n_total, capacity = sc.arrays, sc.n, sc.n * world, int(sc.n * meta.get("capacity_factor", 1.0))
    cap = capacity * meta["pairs_per_particle"]
    sim = apbf_b200.Sim(ctx, sc, capacity=capacity, neighbor_capacity=cap, integrate=True, basic_pbf=meta.get("basic_pbf", not meta["adaptive"]),
                        update_transfers=bool(meta.get("update_transfers")), use_binary_search=(args.search == "binary"),
                        transfers=bool(meta.get("transfers")))

    # host copies of the lists in pinned memory (the e2e leg streams them in every step)
    host = {}
    for name, dt, w in apbf_b200.FIELDS:
        a = np.ascontiguousarray(arrays[name], dtype=dt).reshape(-1, w)
        t = torch.from_numpy(a.view(np.int32) if dt == np.uint32 else a).pin_memory()
        host[name] = t
    out_pos = torch.zeros((capacity, 4), dtype=torch.int32).pin_memory()
    out_kw = torch.zeros((capacity,), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    d2h = n * 16 + n * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sim.upload(host, n=n)
    dom = None
    if slab:
        halo_range = float(sc.arrays["kernel_width"].max()) * (1.5 if meta["adaptive"] else 1.0) * 1.05
        backend = multi_gpu.CudaRankBackend(sim, n, world, rank, halo_range, ghost_capacity=ghost_cap)
        comm = multi_gpu.TorchComm(torch.device("cuda", local_rank))
        dom = multi_gpu.SlabDomain(backend, comm, adaptive=meta["adaptive"], solver_iterations=sc.solver_iterations, integrate=True)

    def step():
        if dom is not None:
            dom.substep()
        else:
            sim.substep(1)

    # ---- device-resident throughput ----------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None   # nvidia-smi needs ~0.2 s to deliver its first sample: start it
    for _ in range(args.warmup):                                # with the warm-up so that it is sampling during the timed steps
        step()
    barrier()
    if dom is not None and dom.timing is not None:
        dom.timing.clear()          # APBF_MG_TIMING: sections of the timed steps only
    launches0 = ctx.launch_count
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ctx.profile(False)
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    prof = ctx.profile_read()
    stats = sim.stats()
    slab_stats = dict(dom.stats, halo_bytes=comm.bytes_sent, messages=comm.messages) if dom is not None else None
    if dom is not None and dom.timing is not None and rank == 0:
        tot = sum(dom.timing.values())
        sys.stderr.write("slab sections (ms/substep, serialised): " + ", ".join(
            f"{k} {v / args.steps * 1e3:.3f}" for k, v in dom.timing.items()) + f" | total {tot / args.steps * 1e3:.3f}\n")
    if meta["adaptive"]:
        # the fused search never builds the unpruned list; one more (untimed) substep counts what it would have held
        ctx.set_search_stats(True)
        step()
        stats = dict(stats, pairs_searched=sim.stats()["pairs_searched"])
        ctx.set_search_stats(False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end to end: host buffers in, host buffers out, every step ---------------------------------------------------
    def e2e_step():
        sim.upload(host, n=n)
        if dom is not None:
            dom.n_own, dom.gid_base = n, 0     # the uploaded brick is the initial one again
            backend.set_counts(n, n, 0)
        step()
        sim.download({"position": out_pos, "kernel_width": out_kw})   # synchronises

    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 20))
    for _ in range(2 if e2e_steps else 0):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    cells = 1 << (sc.res_log2 * sc.dims)
    bits = 4 * -(-(sc.res_log2 * sc.dims + 1) // 4)
    per_launch, substep_bytes = algorithmic_bytes(n, stats["pairs_searched"], stats["pairs_kept"], cells, bits,
                                                 sc.solver_iterations, meta["adaptive"])
    traffic, traffic_src = None, None
    try:   # per-launch DRAM bytes of the passes from the committed ncu --set full capture of this workload (tools/ncu_traffic.py)
        tj = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")))
    except (OSError, ValueError):
        tj = None
    timed = {k: v for k, v in prof.items() if v[1] > 0}
    top = max((k for k in timed if k in per_launch), key=lambda k: timed[k][0])
    top_ms = timed[top][0] / timed[top][1]
    achieved = per_launch[top] / (top_ms * 1e-3) / 1e9
    if tj and world == 1 and top in tj.get("passes", {}):
        traffic, traffic_src = tj["passes"][top]["dram_bytes_per_launch"], tj["source"]
    passes = {k: {"ms_per_launch": round(v[0] / v[1], 4), "launches": v[1], "share": round(v[0] / ms, 4),
                  **({"gbs": round(per_launch[k] / (v[0] / v[1] * 1e-3) / 1e9, 1)} if k in per_launch else {})}
              for k, v in timed.items()}

    value = n_total * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": sc.name, "particles_per_gpu": n, "adaptive_kernel_width": meta["adaptive"],
                   "solver_iterations": sc.solver_iterations, "search": args.search, "res_log2": sc.res_log2,
                   "pairs_searched": stats["pairs_searched"], "pairs_kept": stats["pairs_kept"],
                   "pairs_unmirrored": stats["pairs_unmirrored"],
                   "multi_gpu": "single" if world == 1 else ("one scene in bricks (top key bits), halo exchange over NCCL send/recv" if slab
                                                             else "independent replicas"),
                   **({"particles_total": n_total, "slab": slab_stats} if slab else {}),
                   "l2": "working set (lists + pair list) exceeds the 126 MB L2" if (76 * n + 8 * stats["pairs_searched"]) > 126e6
                         else "working set fits L2; no flush between steps"},
        "gpu_launches": launches,
        "e2e": {"value": n_total * e2e_steps / e2e_s if e2e_steps else None, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": per_launch[top], "ms_per_launch": top_ms,
                     "substep": {"algorithmic_bytes": substep_bytes, "achieved": substep_bytes / (ms / args.steps * 1e-3) / 1e9,
                                 "frac": substep_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}},
        "passes": passes,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        sample = {"dam_break_1M": "dam_break_64k", "uniform_64": "uniform_32", "dam_break_1M_default_mode": "dam_break_64k_default_mode",
              "dam_break_1M_split_merge": "dam_break_64k_split_merge"}.get(args.workload, args.workload)
        threads = os.cpu_count() or 1
        cpu_steps = 30                           # bounded sample: 10-30 s of CPU work
        v, sec, ssc = time_oracle(sample, cpu_steps, 1, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{sample}: {ssc.n} particles, {cpu_steps} substeps of the {args.workload} workload "
                                          f"({sec * cpu_steps:.1f} s of CPU work)"}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--search", default="green", choices=["green", "binary"],
                    help="neighborhood_green (default, pool.cpp NEIGHBORHOOD_TYPE 1) or neighborhood_binary_search (type 3; 1 GPU only)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent copies of the 1-GPU scene instead of one scene in bricks")
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
