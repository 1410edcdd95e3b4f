"""bench.py, N > 1: ONE scene across the GPUs of the node in bricks, one process per GPU.

Main line: the dam break of BASELINE.json configs[1] with 10^6 particles per GPU (weak scaling).  `extra` (unless
--no-mg-extra): the uniform block of configs[4] at 8 M particles per GPU (64 M on 8 GPUs) next to the same per-GPU load on
one GPU, and `n_rank_parity`: a small scene run on N ranks and on one rank must agree bit for bit."""
import os
import sys
import time

import numpy as np


def _brick_arrays(sc, rank, world, multi_gpu):
    owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
    arrays = {k: np.ascontiguousarray(v[owner == rank]) for k, v in sc.arrays.items()}
    arrays["index_list"] = np.arange(len(arrays["position"]), dtype=np.uint32)
    return arrays


class SlabRun:
    """one scene in bricks on this rank: Sim + the slab protocol (library loop by default, Python loop with --mg-python)"""

    def __init__(self, gpu, torch, sc, meta, rank, world, local_rank, python_loop, ghost_frac=None, brick_arrays=None, n_total=None):
        from apbf_b200 import multi_gpu
        self.gpu, self.torch, self.sc, self.meta = gpu, torch, sc, meta
        self.ctx = gpu.Context(device=local_rank, dims=sc.dims)
        self.ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius)
        arrays = brick_arrays if brick_arrays is not None else _brick_arrays(sc, rank, world, multi_gpu)   # (a scene seeded brick by brick)
        self.n = len(arrays["position"])
        self.n_total = int(n_total if n_total is not None else sc.n)
        # ghosts: the layers of particles within the halo range of the brick's faces
        kw_max = float(arrays["kernel_width"].max()) if self.n else float(sc.arrays["kernel_width"].max())
        halo_range = kw_max * (1.5 if meta["adaptive"] else 1.0) * 1.05
        side = max(self.n, 1) ** (1.0 / 3.0)
        layers = halo_range / 2.0 + 1.0
        self.ghost_cap = int(max(400_000 if self.n <= 1_200_000 else 0, 6.0 * side * side * layers * 1.3)) if ghost_frac is None else int(self.n * ghost_frac)
        self.capacity = int(self.n * 1.1) + self.ghost_cap
        self.sim = gpu.Sim(self.ctx, sc, capacity=self.capacity, neighbor_capacity=self.capacity * meta["pairs_per_particle"], integrate=True,
                           basic_pbf=not meta["adaptive"], use_binary_search=(meta.get("search") == "binary"))
        self.sim.upload(arrays, n=self.n)
        self.arrays = arrays
        if python_loop or not hasattr(multi_gpu, "LibraryDomain"):
            backend = multi_gpu.CudaRankBackend(self.sim, self.n, world, rank, halo_range, ghost_capacity=self.ghost_cap)
            comm = multi_gpu.TorchComm(torch.device("cuda", local_rank))
            self.dom = multi_gpu.SlabDomain(backend, comm, adaptive=meta["adaptive"], solver_iterations=sc.solver_iterations, integrate=True)
            self.driver = "python (apbf_b200/multi_gpu.py: SlabDomain over torch.distributed)"
        else:
            self.dom = multi_gpu.LibraryDomain(self.sim, self.n, world, rank, halo_range, ghost_capacity=self.ghost_cap, adaptive=meta["adaptive"],
                                               solver_iterations=sc.solver_iterations)
            how = ("peer to peer: pack kernels store into the receiver's buffer over NVLink and raise a flag, unpack kernels wait on it"
                   if self.dom.transport == "p2p" else "grouped ncclSend / ncclRecv")
            self.driver = f"library (apbf_sim_mg_substep: route, halo, search, solve and their exchanges in C++ on the context's stream; exchanges {how})"

    def step(self):
        self.dom.substep()

    def close(self):
        if hasattr(self.dom, "close"):
            self.dom.close()
        self.sim.close()
        self.ctx.close()
        self.torch.cuda.empty_cache()


def _time_steps(torch, dist, run, steps, warmup):
    for _ in range(warmup):
        run.step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        run.step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def n_rank_parity(gpu, torch, dist, rank, world, local_rank, make_scene, python_loop):
    """the headline's dam break at half the edge length (adaptive; 125 000 particles per rank, balanced over the bricks like the
    full-size scene) and a 48^3 uniform block: N ranks in bricks vs rank 0 alone, 3 substeps, every list compared bit for bit
    after gathering the bricks in rank order"""
    from apbf_b200 import scenes
    import bench
    half = {k: (v // 2 if k in ("nx", "ny", "nz") else v) for k, v in bench.SLAB_DAM_BREAK[world].items() if k != "res_log2"}
    out = {}
    for name, sc, meta in ((f"dam_break_{world}x125k_adaptive", scenes.dam_break(adaptive=True, grid=bench.DAM_BREAK_GRID, **half), dict(adaptive=True, pairs_per_particle=150)),
                           ("uniform_48", scenes.uniform_block(48, jitter=0.1, shuffle=True), dict(adaptive=False, pairs_per_particle=40))):
        run = SlabRun(gpu, torch, sc, meta, rank, world, local_rank, python_loop, ghost_frac=1.5)
        for _ in range(3):
            run.step()
        host = gpu.empty_host_arrays(run.capacity)
        n_own = run.dom.n_owned()
        run.sim.download(host)
        mine = {k: np.ascontiguousarray(host[k][:n_own]).view(np.int32).reshape(n_own, -1) for k in ("position", "velocity", "kernel_width")}
        parts = [None] * world
        dist.all_gather_object(parts, mine)          # (small scenes: through the host is fine)
        counts = [len(p["position"]) for p in parts]
        gathered = {k: np.concatenate([p[k] for p in parts]) for k in mine}
        run.close()
        row = None
        if rank == 0:
            ctx = gpu.Context(device=local_rank, dims=sc.dims)
            ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if meta["adaptive"] else 1, mSmallestTargetRadius=sc.smallest_target_radius)
            sim = gpu.Sim(ctx, sc, neighbor_capacity=sc.n * meta["pairs_per_particle"], integrate=True, basic_pbf=not meta["adaptive"])
            sim.upload(sc.arrays)
            sim.substep(3)
            one = gpu.empty_host_arrays(sc.n)
            sim.download(one)
            sim.close(); ctx.close()
            row = {"particles": int(sc.n), "substeps": 3, "owned_per_rank": [int(c) for c in counts]}
            for k in ("position", "velocity", "kernel_width"):
                a = np.ascontiguousarray(one[k]).view(np.int32).reshape(sc.n, -1)
                row[k + "_bit_exact"] = bool(a.shape == gathered[k].shape and np.array_equal(a, gathered[k]))
            row["ok"] = all(v for kk, v in row.items() if kk.endswith("_bit_exact"))
        out[name] = row
        dist.barrier()
    return out


def run(args, rank, world, local_rank, peak, peak_src, emit, ClockSampler, make_scene, roofline_of, METRIC, UNIT):
    import torch
    import torch.distributed as dist
    import apbf_b200 as gpu

    sc, meta = make_scene(args.workload, world, args.res_log2)
    meta = dict(meta, search=args.search)   # --search binary: the reference's compiled default search over owned particles + ghosts
    if not meta.get("slab"):
        raise SystemExit(f"workload {args.workload} has no brick layout: use --replicas")
    run_ = SlabRun(gpu, torch, sc, meta, rank, world, local_rank, args.mg_python)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        run_.step()
    dist.barrier(); torch.cuda.synchronize()
    ctx = run_.ctx
    launches0 = ctx.launch_count
    ms = _time_steps(torch, dist, run_, args.steps, 0)          # the timed region: nothing but the substeps on the stream
    launches = ctx.launch_count - launches0
    prof_steps = max(3, min(args.steps, 20))                     # per-pass durations: further substeps with events around every pass
    ctx.profile(True)
    prof_ms = _time_steps(torch, dist, run_, prof_steps, 0)
    ctx.profile(False)
    prof = ctx.profile_read()
    stats = run_.sim.stats()
    list_state = ctx.list_state()   # which form of the passes the device selected (ghost queries skipped, one-gather apply sweep, ...)
    if meta["adaptive"]:
        ctx.set_search_stats(True)
        run_.step()
        stats = dict(stats, pairs_searched=run_.sim.stats()["pairs_searched"])
        ctx.set_search_stats(False)
    clocks = sampler.stop() if sampler else None
    slab_stats = dict(run_.dom.stats) if hasattr(run_.dom, "stats") else {}
    flags = ctx.device_flags()

    # ---- end to end: this rank's brick lives in pinned host memory; lists in, one substep, positions and widths out -------------
    e2e = None
    if not args.no_e2e:
        host = {}
        for name, dt, w in gpu.FIELDS:
            a = np.ascontiguousarray(run_.arrays[name], dtype=dt).reshape(-1, w)
            host[name] = torch.from_numpy(a.view(np.int32) if dt == np.uint32 else a).pin_memory()
        out_pos = torch.zeros((run_.capacity, 4), dtype=torch.int32).pin_memory()
        out_kw = torch.zeros((run_.capacity,), dtype=torch.float32).pin_memory()
        n0 = run_.n
        steps_e = max(3, min(args.steps, 10))

        def e2e_step():
            run_.sim.upload(host, n=n0)
            run_.dom.reset(n0)
            run_.step()
            run_.sim.download({"position": out_pos, "kernel_width": out_kw})
        for _ in range(2):
            e2e_step()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            e2e_step()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": sc.n * steps_e / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": sum(x.numel() * x.element_size() for x in host.values()),
               "d2h_bytes_per_step": n0 * 20, "steps": steps_e,
               "what": "per rank: apbf_sim_upload of the brick's lists (pinned host memory) -> one substep in bricks -> download of positions and kernel widths"}
    n_local = run_.n
    driver = run_.driver
    run_.close()

    extra, parity = None, None
    if not args.no_mg_extra:
        parity = n_rank_parity(gpu, torch, dist, rank, world, local_rank, make_scene, args.mg_python)
        # configs[4] at 8 M particles per GPU: uniform_200 on one GPU's worth per rank (uniform_400 = 64 M on 8 GPUs)
        from apbf_b200 import scenes
        side = {2: 252, 4: 318, 8: 400}[world]
        big, big_total = scenes.uniform_block_brick(side, rank, world)       # this rank's brick only: the whole lattice never exists on one host
        bmeta = dict(adaptive=False, pairs_per_particle=40, slab=True)
        brun = SlabRun(gpu, torch, big, bmeta, rank, world, local_rank, args.mg_python, brick_arrays=big.arrays, n_total=big_total)
        bms = _time_steps(torch, dist, brun, 5, 3)
        bstats = brun.sim.stats()
        bflags = brun.ctx.device_flags()
        extra = {f"uniform_{side}_in_bricks": {"particles_total": int(big_total), "particles_per_gpu": int(big_total // world), "res_log2": big.res_log2,
                                               "ms_per_step": bms / 5, "value": big_total * 5 / (bms * 1e-3), "pairs_rank0": bstats["pairs_kept"],
                                               "device_flags": bflags, "ghost_capacity": brun.ghost_cap, "driver": brun.driver[:7],
                                               "slab": dict(brun.dom.stats) if hasattr(brun.dom, "stats") else {},
                                               "compare_with": "extra.uniform_200 of the N = 1 line (8 M particles on one GPU)"}}
        brun.close()
        # configs[3]: the waterfall scene (11 collision boxes), 252^3 = 16 M particles in all, slab-partitioned over the N GPUs
        # (compare with extra.waterfall_16M of the N = 1 line: the same scene on one GPU -- strong scaling)
        wsc = scenes.waterfall()
        wmeta = dict(adaptive=False, pairs_per_particle=40, slab=True)
        wrun = SlabRun(gpu, torch, wsc, wmeta, rank, world, local_rank, args.mg_python)
        wms = _time_steps(torch, dist, wrun, 5, 3)
        wstats = wrun.sim.stats()
        extra["waterfall_16M_in_bricks"] = {"particles_total": int(wsc.n), "particles_this_rank": int(wrun.n), "res_log2": wsc.res_log2, "collision_boxes": int(len(wsc.box_min)),
                                            "ms_per_step": wms / 5, "value": wsc.n * 5 / (wms * 1e-3), "pairs_rank0": wstats["pairs_kept"],
                                            "device_flags": wrun.ctx.device_flags(), "scaling": "strong (one 16 M scene over N GPUs)",
                                            "slab": dict(wrun.dom.stats) if hasattr(wrun.dom, "stats") else {},
                                            "compare_with": "extra.waterfall_16M of the N = 1 line"}
        wrun.close()

    if rank == 0:
        r = dict(sc=sc, meta=meta, n=n_local, ms=ms, steps=args.steps, prof=prof, prof_ms=prof_ms, stats=stats)
        roof, passes = roofline_of(r, peak, args.workload, world)
        roof["peak_source"] = peak_src
        line = {
            "metric": METRIC, "value": sc.n * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": sc.name, "particles_per_gpu": n_local, "particles_total": int(sc.n),
                       "adaptive_kernel_width": meta["adaptive"], "solver_iterations": sc.solver_iterations, "search": args.search, "res_log2": sc.res_log2,
                       "grid_cell": [round((h - l) / (1 << sc.res_log2), 3) for l, h in zip(sc.min_pos, sc.max_pos)],
                       "pairs_searched": stats["pairs_searched"], "pairs_kept": stats["pairs_kept"], "pairs_unmirrored": stats["pairs_unmirrored"],
                       "device_flags": flags, "list_state": list_state, "multi_gpu": "one scene in bricks (top bits of the cell key), ghost particles, halo exchange every solver iteration",
                       "driver": driver, "slab": slab_stats,
                       "l2": "working set (lists + pair list) exceeds the 126 MB L2"},
            "gpu_launches": launches,
            "e2e": e2e if e2e else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 0},
            "roofline": roof, "passes": passes, "clocks": clocks,
        }
        if extra:
            line["extra"] = extra
        if parity:
            line["n_rank_parity"] = parity
        emit(line)
    dist.destroy_process_group()
