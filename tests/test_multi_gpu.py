"""N-rank CUDA run (one process per GPU, NCCL) against the 1-rank CUDA run: the slab protocol must not change a bit.
Needs >= 2 GPUs (skipped otherwise); the host logic itself is covered on CPU by tests/test_multi_rank_gloo.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(adaptive, side, binary=False):
    from apbf_b200 import scenes
    # (binary search: a particle's id is its rank in Morton-code order, which one GPU and a brick number differently, and the wall
    # jitter of box_collision.comp:46-47 hashes the id -- those cases keep the particles out of wall contact)
    sc = scenes.waterdrop(side, jitter=0.1, wall_gap=6.0) if adaptive else scenes.uniform_block(side, jitter=0.2, shuffle=True, wall_gap=6.0 if binary else 0.0)
    sc.arrays["position"][:, 3] = np.arange(sc.n, dtype=np.int32)
    sc.arrays["pos_backup"][:, :3] -= np.array([6000000, 7200000, 9000000], np.int32)  # velocity_handling starts with mLastDeltaTime = 1 (velocity_handling.h:18)
    return sc


def _worker(rank, world, port, adaptive, side, steps, out_dir, default_mode=False, library=False, transport=None, binary=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import apbf_b200
    from apbf_b200 import multi_gpu
    sc = _scene(adaptive or default_mode, side, binary)
    owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
    mine = {k: np.ascontiguousarray(v[owner == rank]) for k, v in sc.arrays.items()}
    n_own = len(mine["position"])
    mine["index_list"] = np.arange(n_own, dtype=np.uint32)
    ctx = apbf_b200.Context(device=rank, dims=sc.dims)
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if adaptive else 1, mSmallestTargetRadius=sc.smallest_target_radius)
    cap = sc.n                      # room for every particle plus ghosts on one rank
    sim = apbf_b200.Sim(ctx, sc, capacity=cap, neighbor_capacity=cap * (700 if adaptive or default_mode else 80), integrate=True,
                        basic_pbf=not (adaptive or default_mode), update_transfers=default_mode, use_binary_search=binary)
    sim.upload(mine, n=n_own)
    halo_range = float(sc.arrays["kernel_width"].max()) * (1.5 if adaptive else 1.0) * 1.05
    if library:   # the whole substep inside the library: apbf_sim_mg_substep
        dom = multi_gpu.LibraryDomain(sim, n_own, world, rank, halo_range, ghost_capacity=cap, adaptive=adaptive, solver_iterations=4, headroom=4.0,
                                         transport=transport)
        assert dom.transport == transport, (dom.transport, transport)   # no silent change of transport
    else:         # the same protocol driven from Python, exchange by exchange
        backend = multi_gpu.CudaRankBackend(sim, n_own, world, rank, halo_range, ghost_capacity=cap)
        dom = multi_gpu.SlabDomain(backend, multi_gpu.TorchComm(torch.device("cuda", rank)), adaptive=adaptive, solver_iterations=4, integrate=True,
                                   update_transfers=default_mode, width_from_boundary_distance=default_mode)
    migrated = 0
    for _ in range(steps):
        dom.substep()
        migrated += dom.stats["migrated"]
    out = apbf_b200.empty_host_arrays(cap)
    n = sim.download(out)
    assert n == dom.n_owned(), (n, dom.n_owned())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), migrated=migrated, ghosts=dom.stats["ghosts"], flags=ctx.device_flags(),
             **{k: v[:n] for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("adaptive,side", [(False, 24), (True, 20)])
def test_n_rank_equals_one_rank_binary_search(tmp_path, adaptive, side):
    """the reference's compiled default search (NEIGHBORHOOD_TYPE 3, neighborhood_binary_search.cpp:22-75) over owned particles +
    ghosts: fixed widths, and adaptive widths (search fused with spread_kernel_width); library loop, peer-to-peer transport"""
    test_n_rank_equals_one_rank(tmp_path, adaptive, side, False, "library_p2p", binary=True)


@pytest.mark.gpu
@pytest.mark.parametrize("driver", ["python", "library_p2p", "library_nccl"])
@pytest.mark.parametrize("adaptive,side,default_mode", [(False, 24, False), (True, 20, False), (False, 20, True)])
def test_n_rank_equals_one_rank(tmp_path, adaptive, side, default_mode, driver, binary=False):
    """driver: the protocol exchange by exchange from Python over torch.distributed; the library's own loop (apbf_sim_mg_substep) with
    the peer-to-peer transport (pack kernels store into the receiver's buffer over NVLink, flags instead of NCCL); the same loop over
    grouped ncclSend / ncclRecv"""
    library = driver != "python"
    transport = {"python": None, "library_p2p": "p2p", "library_nccl": "nccl"}[driver]
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import apbf_b200
    world = 4 if torch.cuda.device_count() >= 4 else 2
    steps = 3
    port = 29500 + (os.getpid() % 1000) + int(adaptive) + 2 * int(default_mode) + 4 * ["python", "library_p2p", "library_nccl"].index(driver)
    mp.spawn(_worker, args=(world, port + (16 if binary else 0), adaptive, side, steps, str(tmp_path), default_mode, library, transport, binary), nprocs=world, join=True)
    sc = _scene(adaptive or default_mode, side, binary)
    ctx = apbf_b200.Context(dims=sc.dims)
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0 if adaptive else 1, mSmallestTargetRadius=sc.smallest_target_radius)
    sim = apbf_b200.Sim(ctx, sc, neighbor_capacity=sc.n * (700 if adaptive or default_mode else 80), integrate=True,
                        basic_pbf=not (adaptive or default_mode), update_transfers=default_mode, use_binary_search=binary)
    sim.upload(sc.arrays)
    sim.substep(steps)
    exp = apbf_b200.empty_host_arrays(sc.n)
    assert sim.download(exp) == sc.n
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    assert all(int(p["flags"]) == 0 for p in parts)
    pos = np.concatenate([p["position"] for p in parts])
    assert len(pos) == sc.n and len(set(pos[:, 3].tolist())) == sc.n
    assert sum(int(p["migrated"]) for p in parts) > 0 and all(int(p["ghosts"]) > 0 for p in parts)
    got, ref = np.argsort(pos[:, 3]), np.argsort(exp["position"][:, 3])
    for k in ("position", "velocity", "kernel_width", "boundariness") + (("boundary_distance", "target_radius") if default_mode else ()):
        a = np.concatenate([p[k] for p in parts])[got]
        assert np.array_equal(a, exp[k][ref]), k      # same kernels, integer accumulators: bit for bit, walls included (global ids)
