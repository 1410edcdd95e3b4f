"""world_size-2 (and 4) run of the slab protocol on CPU: apbf_b200.multi_gpu.SlabDomain over gloo with the numpy/oracle
rank backend (tests/mg_oracle_backend.py).  The N-rank result must equal the 1-rank oracle run bit for bit (matched by a
tag in position.w): routing, halo lists, ghost slots after the sort and the order of the exchanges are what is tested."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene(adaptive):
    from apbf_b200 import scenes
    if adaptive:
        sc = scenes.waterdrop(14, jitter=0.1, wall_gap=6.0)
    else:
        sc = scenes.uniform_block(12, jitter=0.2, shuffle=True, wall_gap=6.0)
    sc.arrays["position"][:, 3] = np.arange(sc.n, dtype=np.int32)
    # give the block a drift so that particles cross the brick boundaries within a few substeps
    sc.arrays["pos_backup"][:, :3] -= np.array([100000, 120000, 150000], np.int32)
    return sc


def _settings(orc, sc, adaptive):
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0 if adaptive else 1
    s.mSmallestTargetRadius = sc.smallest_target_radius
    return s


def _worker(rank, world, port, adaptive, steps, out_dir, default_mode=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from apbf_b200 import multi_gpu
    from oracle import oracle as orc
    from mg_oracle_backend import OracleRankBackend
    orc.set_threads(1)
    sc = _scene(adaptive or default_mode)
    s = _settings(orc, sc, adaptive)
    owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
    mine = {k: v[owner == rank] for k, v in sc.arrays.items() if k != "index_list"}
    halo_range = float(sc.arrays["kernel_width"].max()) * (1.5 if adaptive else 1.0) * 1.05
    backend = OracleRankBackend(mine, sc, s, rank, world, halo_range, adaptive, cap_pairs=sc.n * 700)
    dom = multi_gpu.SlabDomain(backend, multi_gpu.TorchComm(), adaptive=adaptive, solver_iterations=4, integrate=True,
                               update_transfers=default_mode, width_from_boundary_distance=default_mode)
    migrated = 0
    for _ in range(steps):
        dom.substep()
        migrated += dom.stats["migrated"]
    a = backend.owned_arrays()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), migrated=migrated, ghosts=dom.stats["ghosts"], **a)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,adaptive,default_mode", [(2, False, False), (2, True, False), (4, False, False), (2, False, True)])
def test_slab_protocol_equals_single_rank(orc, tmp_path, world, adaptive, default_mode):
    """default_mode: the reference's default adaptive mode -- kernel width from the boundary distance before the search and
    update_transfers after the solver (its flood fill crosses the bricks)"""
    steps = 3
    port = 29000 + (os.getpid() % 2000) + world * 3 + int(adaptive) + 7 * int(default_mode)
    mp.spawn(_worker, args=(world, port, adaptive, steps, str(tmp_path), default_mode), nprocs=world, join=True)
    # single rank: the plain oracle substep
    sc = _scene(adaptive or default_mode)
    s = _settings(orc, sc, adaptive)
    orc.set_threads(1)
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    for _ in range(steps):
        orc.substep(st, s, dims=sc.dims, basic_pbf=not (adaptive or default_mode), solver_iterations=4, min_pos=sc.min_pos, max_pos=sc.max_pos,
                    res_log2=sc.res_log2, box_min4=sc.box_min, box_max4=sc.box_max, cap=sc.n * 700, integrate=True,
                    update_transfers=default_mode)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    pos = np.concatenate([p["position"] for p in parts])
    kw = np.concatenate([p["kernel_width"] for p in parts])
    vel = np.concatenate([p["velocity"] for p in parts])
    assert len(pos) == sc.n and len(set(pos[:, 3].tolist())) == sc.n            # nobody lost, nobody duplicated
    assert sum(int(p["migrated"]) for p in parts) > 0                           # particles really changed owner
    assert all(int(p["ghosts"]) > 0 for p in parts)
    got, exp = np.argsort(pos[:, 3]), np.argsort(st.position[:, 3])
    assert np.array_equal(pos[got], st.position[exp])                           # bit for bit
    assert np.array_equal(kw[got], st.kernel_width[exp])
    assert np.array_equal(vel[got], st.velocity[exp])
    if default_mode:
        for k in ("boundary_distance", "target_radius", "boundariness"):
            a = np.concatenate([p[k] for p in parts])
            assert np.array_equal(a[got], getattr(st, k)[exp]), k
        assert len(np.unique(st.boundary_distance)) > 50
