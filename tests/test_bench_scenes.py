"""bench.py's workloads (CPU-only checks): the N-GPU dam-break scenes must split evenly over the bricks, and the
algorithmic-byte model must follow SURVEY.md 8(d)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from apbf_b200 import multi_gpu, scenes  # noqa: E402


@pytest.mark.parametrize("grid", ["pool", "cube"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_dam_break_is_balanced(world, grid):
    """same construction as bench.SLAB_DAM_BREAK at a tenth of the edge length: every brick owns the same number of particles
    (the bricks are the halves of the grid along z, then y, then x: the top bits of the cell key), with the pool-hugging grid and
    with the equal-extent grid bench.py times (every axis keeps its centre, so the cuts stay where they were)"""
    args = dict(bench.SLAB_DAM_BREAK[world])
    for k in ("nx", "ny", "nz"):
        args[k] //= 5
    args.pop("res_log2", None)
    sc = scenes.dam_break(adaptive=True, grid=grid, **args)
    owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
    counts = np.bincount(owner, minlength=world)
    assert counts.sum() == sc.n and counts.min() > 0
    assert counts.max() - counts.min() <= 0.02 * sc.n / world, counts


def test_timed_dam_break_searches_on_cubic_cells():
    """bench.py's dam breaks: equal extents on all axes around the same centres as the pool-hugging grid, cell edge within a
    factor 1.5 of 1.1 kernel widths at every GPU count; the particles themselves do not depend on the grid"""
    assert bench.DAM_BREAK_GRID == "cube"
    for world in (1, 2, 4, 8):
        args = dict(bench.SLAB_DAM_BREAK.get(world, dict(nx=100, ny=100, nz=100)))
        for k in ("nx", "ny", "nz"):
            args[k] //= 10
        pool = scenes.dam_break(adaptive=True, grid="pool", **args)
        cube = scenes.dam_break(adaptive=True, grid="cube", **args)
        ext = [h - l for l, h in zip(cube.min_pos, cube.max_pos)]
        assert max(ext) - min(ext) < 1e-3 * max(ext)
        assert np.allclose([0.5 * (l + h) for l, h in zip(cube.min_pos, cube.max_pos)], [0.5 * (l + h) for l, h in zip(pool.min_pos, pool.max_pos)])
        assert all(cl <= pl + 1e-4 and ch >= ph - 1e-4 for cl, ch, pl, ph in zip(cube.min_pos, cube.max_pos, pool.min_pos, pool.max_pos))
        assert all(np.array_equal(cube.arrays[k], pool.arrays[k]) for k in pool.arrays)
    # the resolution rule at the full-size extents (620, 620, 820, 1220 r)
    for extent, res in ((620.0, 7), (820.0, 8), (1220.0, 8)):
        assert scenes._res_near(extent, 4.4) == res and 4.4 / 1.5 < extent / (1 << res) < 4.4 * 1.5


def test_full_size_slab_scenes_have_a_million_particles_per_gpu():
    for world, a in bench.SLAB_DAM_BREAK.items():
        assert a["nx"] * a["ny"] * a["nz"] * a.get("blocks", 1) == world * 1_000_000


def test_algorithmic_bytes_follow_the_survey():
    n, p, pk, cells, bits, iters = 1000, 30000, 9000, 4096, 16, 4
    per, sub = bench.algorithmic_bytes(n, p, pk, cells, bits, iters, adaptive=True)
    passes = 2
    assert per["hash_sort"] == 20 * n + 16 * n * passes and per["reorder"] == 152 * n and per["cell_ranges"] == 4 * n + 8 * cells
    assert per["density_lambda"] == 8 * pk + 56 * n and per["apply_delta"] == 8 * pk + 40 * n and per["box_collision"] == 36 * n
    assert sub == (196 + 16 * passes) * n + 16 * cells + 8 * p + iters * (16 * pk + 132 * n) + 8 * p + 8 * pk + 12 * n


def test_waterfall_scene_has_eleven_boxes_and_keeps_its_fluid(orc):
    """configs[3] (source/waterfall.cpp:28-48): 4 + 3 walls in the plane, 2 + 2 in z; the closed top pool holds the block through
    oracle substeps with the integrator on (box_collision against all eleven boxes every iteration)"""
    sc = scenes.waterfall(10, 10, 10)
    assert sc.box_min.shape == (11, 4) and np.all(sc.box_max[:, :3] > sc.box_min[:, :3])
    shift = sc.box_min[4, :3] - sc.box_min[0, :3]                      # bottom pool = top pool moved down and to the left
    assert shift[0] < 0 and shift[1] < 0 and shift[2] == 0
    assert np.allclose(sc.box_max[4, :3] - sc.box_max[0, :3], shift) and np.allclose(sc.box_min[9, :3] - sc.box_min[7, :3], shift)
    assert scenes.waterfall_boxes((0, 0, 0), (20, 20, 20), 1.0, dims=2)[0].shape == (7, 4)
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    s = orc.default_settings()
    for _ in range(5):
        orc.substep(st, s, dims=3, basic_pbf=True, solver_iterations=4, min_pos=sc.min_pos, max_pos=sc.max_pos, res_log2=sc.res_log2,
                    box_min4=sc.box_min, box_max4=sc.box_max, cap=sc.n * 80, integrate=True)
    pos = st.position[:, :3].astype(np.float64) / 262144.0
    inner_lo, inner_hi = sc.box_max[0, 0], sc.box_min[2, 0]             # between the left and the right wall of the top pool
    assert np.all(pos[:, 0] > inner_lo - 1e-3) and np.all(pos[:, 0] < inner_hi + 1e-3)
    assert np.all(pos[:, 1] > sc.box_max[1, 1] - 1e-3) and np.all(pos > np.asarray(sc.min_pos)) and np.all(pos < np.asarray(sc.max_pos))


def test_waterfall_full_size():
    assert 252 ** 3 == 16_003_008 and bench.make_scene("waterfall_64k")[0].n == 64000


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the oracle on the host cores, kind "port"): one JSON line on stdout with the keys
    the driver reads"""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "uniform_32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["config"]["workload"] == "uniform_32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
