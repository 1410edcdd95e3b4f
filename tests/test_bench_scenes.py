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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_dam_break_is_balanced(world):
    """same construction as bench.SLAB_DAM_BREAK at a tenth of the edge length: every brick owns the same number of particles
    (the bricks are the halves of the grid along z, then y, then x: the top bits of the cell key)"""
    args = dict(bench.SLAB_DAM_BREAK[world])
    for k in ("nx", "ny", "nz"):
        args[k] //= 5
    args.pop("res_log2", None)
    sc = scenes.dam_break(adaptive=True, **args)
    owner = multi_gpu.owner_rank_of_positions(sc.arrays["position"], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims, world)
    counts = np.bincount(owner, minlength=world)
    assert counts.sum() == sc.n and counts.min() > 0
    assert counts.max() - counts.min() <= 0.02 * sc.n / world, counts


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
