#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

Two kinds:
  kats.json     -- the reference's OWN known-answer vectors for this path, copied as data from its self tests
                   (source/test.cpp:91-105 apply_edit, :186-197 prefix sum, :284-305 / :343-364 / :402-425 sort,
                   :117-184 hidden edits, :487-513 duplicate_these, :427-449 delete_these, :623 Z-curve cells) and from the
                   shader text (calculate_position_code.comp:28-30).  These pin the oracle AND the CUDA path.
  *.npz         -- outputs of the CPU oracle (oracle/apbf_oracle.c, single-threaded) on small seeded scenes for the passes
                   the reference has no enabled test for (neighbour search, incompressibility, kernel width, box collision,
                   whole substeps).  The reference cannot be built or run in this image (Vulkan/GLSL, MSVC), so these are NOT
                   reference outputs: they freeze the oracle so that neither it nor the kernels can drift unnoticed.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from apbf_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402

KATS = {
    "apply_edit": {"list": [77, 3, 9999, 4294967295, 0], "edit": [1, 3, 1, 4], "expected": [3, 4294967295, 3, 0], "ref": "source/test.cpp:91-105"},
    "prefix_sum": {"values": [43, 1, 4567, 0, 1, 0, 84523487], "expected": [43, 44, 4611, 4611, 4612, 4612, 84528099], "ref": "source/test.cpp:186-197"},
    "sort": {"keys": [15, 2, 1234, 2, 0, 4294967295, 1, 4294967294], "expected_payload": [4, 6, 1, 3, 0, 2, 7, 5], "ref": "source/test.cpp:284-305,402-425"},
    "sort_small_values": {"keys": [15, 2, 3, 2, 0, 14, 1, 14], "expected_payload": [4, 6, 1, 3, 2, 5, 7, 0], "ref": "source/test.cpp:343-364"},
    "hidden_edit_2": {"edit": [0, 1, 3, 4], "index_a": [0, 1, 2, 3, 4], "index_b": [0, 3, 1], "expected_a": [0, 1, 2, 3], "expected_b": [0, 1, 2],
                      "ref": "source/test.cpp:140-162"},
    "hidden_edit_3": {"edit": [2, 1, 2, 4, 1], "index_a": [0, 1, 2, 3, 4], "index_b": [4, 2, 4], "expected_a": [0, 1, 2, 3, 4], "expected_b": [0, 2, 3, 3],
                      "ref": "source/test.cpp:164-184"},
    "zcurve_cells_6bit_3d": {"cells": [[2, 5, 1], [7, 0, 0], [0, 1, 0], [63, 0, 62]], "expected": [142, 73, 2, 187241], "ref": "source/test.cpp:623"},
    "position_code": {"position": [1, 2, 4], "section0": 273, "ref": "shaders/calculate_position_code.comp:28-30"},
    "three_particles": {"positions": [[0, 0, 0], [1, 0, 0], [0, 1, 0]], "ranges": [1, 1, 2], "expected_pairs": [[0, 1], [0, 2], [1, 0], [2, 0], [2, 1]],
                        "ref": "source/test.cpp:560-563 (set-up of the disabled neighbour test; pairs derived by hand)"},
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_case(name, sc, scale, adaptive, hk=1, gk=1, method=2):
    s = orc.default_settings()
    s.mHeightKernelId, s.mGradientKernelId, s.mBoundarinessCalculationMethod = hk, gk, method
    s.mBaseKernelWidthOnBoundaryDistance = 0 if adaptive else 1
    s.mSmallestTargetRadius = sc.smallest_target_radius
    cap = sc.n * (700 if adaptive else 80)
    orc.set_threads(1)
    out = {}
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    pairs, aux = orc.green_apply(st, s, sc.dims, scale, sc.min_pos, sc.max_pos, sc.res_log2, cap, want_aux=True)
    out.update(sorted_hash=aux["sorted_hash"], sorted_index=aux["sorted_index"], cell_start=aux["cell_start"], cell_end=aux["cell_end"],
               pairs=pairs, position_sorted=st.position.copy())
    if adaptive:
        kept, kwfx = orc.spread_kernel_width_apply(st, s, pairs)
        out.update(kept_pairs=kept, kw_fixed=kwfx, kernel_width=st.kernel_width.copy())
        pairs = kept
    # update_transfers (merge / split off) and the kernel width of the default adaptive mode, on a copy of the searched state:
    # the incompressibility outputs below stay what they were before these two were added
    st2 = st.copy()
    nearest = orc.update_transfers_apply(st2, s, pairs)
    out.update(ut_nearest=nearest, ut_boundary_distance=st2.boundary_distance.copy(), ut_target_radius=st2.target_radius.copy(),
               ut_boundariness=st2.boundariness.copy())
    orc.kernel_width_from_boundary_distance(st2, s)
    out.update(ut_kernel_width=st2.kernel_width.copy())
    a = orc.incompressibility_apply(st, s, sc.dims, pairs, want_aux=True)
    out.update(density=a["density"], grad_sum=a["grad_sum"], sq_grad_sum=a["sq_grad_sum"], lam=a["lam"], position_after=st.position.copy(),
               boundariness=st.boundariness.copy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return {k: sha(v) for k, v in out.items()}


def transfers_case(name="transfers14"):
    """split / merge proper (update_transfers.cpp:14-70, particle_transfer.cpp:10-28): a searched waterdrop(14) with radii, target
    radii and transferring flags drawn so that merges, splits and conflicts all occur; frozen after update_transfers and after
    each of two particle_transfer steps (merges take two steps, splits finish at once)."""
    sc = scenes.waterdrop(14, jitter=0.1)
    s = orc.default_settings()
    s.mMerge, s.mSplit, s.mUpdateTargetRadius, s.mMergeDuration = 1, 1, 0, 2.0 / 60.0
    orc.set_threads(1)
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    pairs = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 300)
    rng = np.random.default_rng(11)
    radius = rng.uniform(0.8, 3.2, sc.n).astype(np.float32)
    radius[rng.random(sc.n) < 0.3] = np.float32(1.5)
    inputs = dict(radius=radius, inverse_mass=(1.0 / (2.0 * radius) ** 3).astype(np.float32),
                  target_radius=rng.uniform(0.7, 4.5, sc.n).astype(np.float32), transferring=(rng.random(sc.n) < 0.1).astype(np.uint32))
    for k, v in inputs.items():
        getattr(st, k)[:] = v
    out = {"in_" + k: v for k, v in inputs.items()}
    cap = 2 * sc.n
    T = orc.Transfers(cap)

    def freeze(tag):
        for k, _, _ in orc.State.FIELDS:
            out[f"{tag}_{k}"] = getattr(st, k).copy()
        out[f"{tag}_rows_source"], out[f"{tag}_rows_target"], out[f"{tag}_rows_time_left"] = T.rows()

    out["nearest"] = orc.update_transfers_full(st, s, 3, pairs, T, hidden_cap=cap, split_duration=0.0)
    freeze("u")
    for step in (1, 2):
        orc.particle_transfer_apply(st, T, 3, float(np.float32(1.0 / 60.0)))
        freeze(f"p{step}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return dict(n=sc.n, capacity=cap, sha256={k: sha(v) for k, v in out.items()})


def main():
    if sys.argv[1:] == ["transfers"]:           # only the split / merge fixture (the others stay byte for byte what they are)
        index = json.load(open(os.path.join(HERE, "index.json")))
        index["transfers14"] = transfers_case()
        with open(os.path.join(HERE, "index.json"), "w") as f:
            json.dump(index, f, indent=1)
        print("wrote transfers14.npz")
        return
    index = {"kats": "kats.json", "scenes": {}}
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
    cases = {
        "block12_gauss": (scenes.uniform_block(12, jitter=0.2, shuffle=True), 1.0, False, 1, 1, 2),
        "block12_cubic_spiky": (scenes.uniform_block(12, jitter=0.2, shuffle=True), 1.0, False, 0, 2, 0),
        "block2d_48": (scenes.uniform_block(48, jitter=0.2, dims=2, shuffle=True), 1.0, False, 1, 1, 2),
        "waterdrop16_adaptive": (scenes.waterdrop(16, jitter=0.1), 1.5, True, 1, 1, 2),
    }
    for name, (sc, scale, adaptive, hk, gk, method) in cases.items():
        index["scenes"][name] = dict(scale=scale, adaptive=adaptive, kernels=[hk, gk], method=method, n=sc.n, sha256=scene_case(name, sc, scale, adaptive, hk, gk, method))
    index["transfers14"] = transfers_case()
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)
    print("wrote", ", ".join(sorted(os.listdir(HERE))))


if __name__ == "__main__":
    main()
