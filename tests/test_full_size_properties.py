"""GPU checks at BASELINE.json's FULL sizes (run with -m gpu), where the CPU oracle would need minutes: properties that do not depend
on the size and can be verified independently of the oracle.

  * configs[4] / maximum sizes: a jitter-free 256^3 lattice (16 777 216 particles) has a closed-form pair count -- every lattice
    offset v with |v| <= 2 spacings contributes prod(n - |v_i|) ordered pairs -- and 32 neighbours for every interior particle;
    the keys come out sorted, the cell tables partition them, the list is grouped by id, free of self pairs and symmetric.
  * configs[1]: the dam break with 10^6 particles, adaptive widths, fused search + spread: the neighbour SETS of 3000 sampled
    particles equal what a k-d tree (scipy, double precision) finds within the prune cutoff, up to pairs within 1e-5 of it; the list
    is symmetric (equal widths), two runs give identical bits, and two substeps keep every particle inside the pool.
"""
import itertools

import numpy as np
import pytest

from apbf_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import apbf_b200
    return apbf_b200


def _pair_keys(p):
    return (p[:, 0].astype(np.uint64) << np.uint64(32)) | p[:, 1].astype(np.uint64)


def test_uniform_256_lattice_closed_form(gpu):
    side = 256
    sc = scenes.uniform_block(side, jitter=0.0, shuffle=True)
    assert sc.n == 16_777_216
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=sc.n * 33)
    dbg = gpu.neighborhood_green(ctx).set_data(L).set_range_scale(1.0).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply(debug=True)
    assert ctx.device_flags() == 0
    # closed form: range = kernel width = 4 r = 2 lattice spacings; offsets with |v|^2 <= 4 (the lattice is exact in fixed point)
    expected = 0
    for v in itertools.product(range(-2, 3), repeat=3):
        if v != (0, 0, 0) and v[0] ** 2 + v[1] ** 2 + v[2] ** 2 <= 4:
            expected += (side - abs(v[0])) * (side - abs(v[1])) * (side - abs(v[2]))
    assert L.pair_count() == expected
    keys = dbg["sorted_key"]
    assert np.all(keys[1:] >= keys[:-1])                                       # sortedness
    cs, ce = dbg["cell_start"].astype(np.int64), dbg["cell_end"].astype(np.int64)
    occ = ce > cs
    assert int((ce - cs)[occ].sum()) == sc.n                                   # the cell tables partition the sorted list ...
    first = np.flatnonzero(occ)
    assert np.array_equal(keys[cs[first]], first.astype(np.uint32)) and np.array_equal(keys[ce[first] - 1], first.astype(np.uint32))   # ... by key
    off = dbg["pair_offsets"].astype(np.int64)
    cnt = np.diff(off)
    assert off[0] == 0 and off[-1] == expected and cnt.max() == 32             # interior particles: 32 neighbours
    assert int((cnt == 32).sum()) == (side - 4) ** 3
    # the 5.3e8 pairs stay on the device for the list checks (torch is plumbing here, as everywhere)
    import torch
    pairs = L.pairs[:expected].view(-1, 2).to(torch.int64)
    ids = torch.repeat_interleave(torch.arange(sc.n, device=pairs.device), torch.from_numpy(cnt).to(pairs.device))
    assert bool(torch.equal(pairs[:, 0], ids))                                 # grouped by id, ascending
    del ids
    assert not bool((pairs[:, 0] == pairs[:, 1]).any())
    k = torch.sort((pairs[:, 0] << 32) | pairs[:, 1]).values
    assert bool((k[1:] != k[:-1]).all())
    m = torch.sort((pairs[:, 1] << 32) | pairs[:, 0]).values
    assert bool(torch.equal(k, m))                                             # symmetric
    del pairs, k, m
    del L
    ctx.close()


def test_dam_break_1M_neighbour_sets_against_kdtree(gpu):
    from scipy.spatial import cKDTree
    sc = scenes.dam_break(100, 100, 100, adaptive=True)
    assert sc.n == 1_000_000
    out = []
    for run in range(2):
        ctx = gpu.Context()
        ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0, mSmallestTargetRadius=sc.smallest_target_radius)
        L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=sc.n * 60)
        op = gpu.neighborhood_green_spread(ctx).set_data(L).set_range_scale(1.5).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2)
        op.apply()
        assert ctx.device_flags() == 0
        pairs = L.read_pairs()
        pos = L.read("position")[:, :3].astype(np.float64) / 262144.0
        kw = L.read("kernel_width")
        out.append((pairs, pos, kw))
        del L
        ctx.close()
    (pairs, pos, kw), (pairs2, pos2, kw2) = out
    assert np.array_equal(pairs, pairs2) and np.array_equal(pos, pos2) and np.array_equal(kw, kw2)   # deterministic to the bit
    assert np.all(kw == np.float32(4.0))                        # equal radii: nothing spreads, the prune cutoff is the width itself
    assert np.all(np.diff(pairs[:, 0].astype(np.int64)) >= 0) and not np.any(pairs[:, 0] == pairs[:, 1])
    k = _pair_keys(pairs)
    assert np.array_equal(np.sort(k), np.sort(_pair_keys(pairs[:, ::-1])))
    # independent neighbour sets for a sample
    tree = cKDTree(pos)
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(sc.n, 3000, replace=False))
    starts = np.searchsorted(pairs[:, 0], sample, side="left")
    ends = np.searchsorted(pairs[:, 0], sample, side="right")
    cut = 4.0
    inner = tree.query_ball_point(pos[sample], cut * (1 - 1e-5))
    outer = tree.query_ball_point(pos[sample], cut * (1 + 1e-5))
    for i, a in enumerate(sample):
        got = set(pairs[starts[i]:ends[i], 1].tolist())
        assert set(inner[i]) - {a} <= got <= set(outer[i]) - {a}, a


def test_dam_break_1M_substeps_stay_in_the_pool(gpu):
    sc = scenes.dam_break(100, 100, 100, adaptive=True)
    ctx = gpu.Context()
    ctx.set_settings(mBaseKernelWidthOnBoundaryDistance=0, mSmallestTargetRadius=sc.smallest_target_radius)
    sim = gpu.Sim(ctx, sc, neighbor_capacity=sc.n * 60, integrate=True, basic_pbf=False)
    sim.upload(sc.arrays)
    sim.substep(3)
    host = gpu.empty_host_arrays(sc.n)
    assert sim.download(host) == sc.n
    assert ctx.device_flags() == 0
    pos = host["position"][:, :3].astype(np.float64) / 262144.0
    assert np.isfinite(pos).all()
    # pool walls (pool.cpp:30-40): left / floor / back walls end at 0, the others start at the pool's size (600 x 300 (+ 2 r) x 200)
    assert pos.min() > -1.0 and pos[:, 0].max() < 601.0 and pos[:, 1].max() < 303.0 and pos[:, 2].max() < 201.0
    assert np.array_equal(np.sort(host["index_list"]), np.arange(sc.n, dtype=np.uint32))
    assert len(np.unique(host["inverse_mass"])) == 1
    sim.close()
    ctx.close()
