"""Committed golden fixtures (tests/golden/, generator: tests/golden/make_golden.py).

  * kats.json: the reference's own known-answer vectors for this path (source/test.cpp, shader text) -- they pin the
    oracle (CPU leg) and the CUDA path through the C-ABI (GPU leg).
  * *.npz: frozen oracle outputs for the passes the reference has no enabled test for.  CPU leg: the oracle still
    reproduces them bit for bit.  GPU leg: the CUDA path matches them -- bit-exact for keys, orders, cell ranges and pair
    lists; within the stated tolerance (units of 2^-18, 1e-5 relative) for the PBF quantities.
"""
import json
import os

import numpy as np
import pytest

from apbf_b200 import scenes

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(HERE, "kats.json")))
INDEX = json.load(open(os.path.join(HERE, "index.json")))


def _scene(name):
    if name.startswith("block12"):
        return scenes.uniform_block(12, jitter=0.2, shuffle=True)
    if name == "block2d_48":
        return scenes.uniform_block(48, jitter=0.2, dims=2, shuffle=True)
    if name == "waterdrop16_adaptive":
        return scenes.waterdrop(16, jitter=0.1)
    raise KeyError(name)


def _settings(orc, meta, sc):
    s = orc.default_settings()
    s.mHeightKernelId, s.mGradientKernelId = meta["kernels"]
    s.mBoundarinessCalculationMethod = meta["method"]
    s.mBaseKernelWidthOnBoundaryDistance = 0 if meta["adaptive"] else 1
    s.mSmallestTargetRadius = sc.smallest_target_radius
    return s


# ---- CPU leg: the oracle against the reference's vectors and against its own frozen outputs -----------------------------------
def test_oracle_matches_reference_kats(orc):
    k = KATS["apply_edit"]
    assert orc.apply_edit(np.array(k["list"], np.uint32), np.array(k["edit"], np.uint32)).tolist() == k["expected"]
    k = KATS["prefix_sum"]
    assert orc.prefix_sum(np.array(k["values"], np.uint32)).tolist() == k["expected"]
    for name in ("sort", "sort_small_values"):
        k = KATS[name]
        keys, vals = orc.sort(np.array(k["keys"], np.uint32), np.arange(len(k["keys"]), dtype=np.uint32))
        assert vals.tolist() == k["expected_payload"] and keys.tolist() == sorted(k["keys"])
    k = KATS["zcurve_cells_6bit_3d"]  # cell (x,y,z) of a [0,64)^3 grid with 6 bits per axis: position = cell + 0.5
    pos = np.zeros((len(k["cells"]), 4), np.int32)
    pos[:, :3] = ((np.array(k["cells"], np.float32) + 0.5) * 262144.0).astype(np.int32)
    assert orc.position_hash(pos, (0, 0, 0), (64, 64, 64), 6, 3).tolist() == k["expected"]
    k = KATS["position_code"]
    pos = np.array([k["position"] + [0]], np.int32)
    assert int(orc.position_code(np.array([0], np.uint32), pos, 0)[0]) == k["section0"]
    k = KATS["three_particles"]
    pos = np.zeros((3, 4), np.int32)
    pos[:, :3] = (np.array(k["positions"], np.float32) * 262144.0).astype(np.int32)
    pairs = orc.brute_force_pairs(np.arange(3, dtype=np.uint32), pos, np.array(k["ranges"], np.float32), 1.0, 64)
    assert sorted(map(tuple, pairs.tolist())) == sorted(map(tuple, k["expected_pairs"]))


@pytest.mark.parametrize("name", sorted(INDEX["scenes"]))
def test_oracle_reproduces_frozen_outputs(orc, name):
    meta = INDEX["scenes"][name]
    sc = _scene(name)
    assert sc.n == meta["n"]
    g = np.load(os.path.join(HERE, name + ".npz"))
    s = _settings(orc, meta, sc)
    orc.set_threads(2)  # integer accumulators: the thread count must not matter
    st = orc.State(**{k: v.copy() for k, v in sc.arrays.items()})
    cap = sc.n * (700 if meta["adaptive"] else 80)
    pairs, aux = orc.green_apply(st, s, sc.dims, meta["scale"], sc.min_pos, sc.max_pos, sc.res_log2, cap, want_aux=True)
    for k in ("sorted_hash", "sorted_index", "cell_start", "cell_end"):
        assert np.array_equal(aux[k], g[k]), k
    assert np.array_equal(pairs, g["pairs"]) and np.array_equal(st.position, g["position_sorted"])
    if meta["adaptive"]:
        pairs, kwfx = orc.spread_kernel_width_apply(st, s, pairs)
        assert np.array_equal(pairs, g["kept_pairs"]) and np.array_equal(kwfx, g["kw_fixed"]) and np.array_equal(st.kernel_width, g["kernel_width"])
    st2 = st.copy()                                   # update_transfers + pool.cpp:77-80 on the searched state
    assert np.array_equal(orc.update_transfers_apply(st2, s, pairs), g["ut_nearest"])
    for k in ("boundary_distance", "target_radius", "boundariness"):
        assert np.array_equal(getattr(st2, k), g["ut_" + k]), k
    orc.kernel_width_from_boundary_distance(st2, s)
    assert np.array_equal(st2.kernel_width, g["ut_kernel_width"])
    a = orc.incompressibility_apply(st, s, sc.dims, pairs, want_aux=True)
    for k in ("density", "grad_sum", "sq_grad_sum", "lam"):
        assert np.array_equal(a[k], g[k]), k
    assert np.array_equal(st.position, g["position_after"]) and np.array_equal(st.boundariness, g["boundariness"])


# ---- GPU leg: the CUDA path through the C-ABI against the same fixtures ------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import apbf_b200
    return apbf_b200


def _dev(a):
    import torch
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()


@pytest.mark.gpu
def test_cuda_matches_reference_kats(gpu):
    import torch
    ctx = gpu.Context()
    lib, h = ctx.lib, ctx.handle
    u32 = lambda t: t.cpu().numpy().view(np.uint32)
    k = KATS["apply_edit"]
    src, edit = _dev(np.array(k["list"], np.uint32)), _dev(np.array(k["edit"], np.uint32))
    dst = torch.zeros(len(k["edit"]), dtype=torch.int32, device="cuda")
    ln = torch.tensor([len(k["edit"])], dtype=torch.int32, device="cuda")
    assert lib.apbf_copy_scattered_read(h, src.data_ptr(), dst.data_ptr(), edit.data_ptr(), ln.data_ptr(), len(k["edit"]), 4) == 0
    assert u32(dst).tolist() == k["expected"]
    k = KATS["prefix_sum"]
    v = _dev(np.array(k["values"], np.uint32))
    ln = torch.tensor([len(k["values"])], dtype=torch.int32, device="cuda")
    gpu.algorithms(ctx).prefix_sum(v, ln, len(k["values"]))
    assert u32(v).tolist() == k["expected"]
    for name in ("sort", "sort_small_values"):
        k = KATS[name]
        n = len(k["keys"])
        keys, vals = _dev(np.array(k["keys"], np.uint32)), _dev(np.arange(n, dtype=np.uint32))
        ok, ov = torch.zeros_like(keys), torch.zeros_like(vals)
        gpu.algorithms(ctx).sort(keys, vals, torch.tensor([n], dtype=torch.int32, device="cuda"), n, ok, ov)
        assert u32(ov).tolist() == k["expected_payload"] and u32(ok).tolist() == sorted(k["keys"])
    for name in ("hidden_edit_2", "hidden_edit_3"):
        k = KATS[name]
        edit = _dev(np.array(k["edit"], np.uint32))
        le = torch.tensor([len(k["edit"])], dtype=torch.int32, device="cuda")
        for idx_key, exp_key in (("index_a", "expected_a"), ("index_b", "expected_b")):
            idx = torch.zeros(8, dtype=torch.int32, device="cuda")
            idx[: len(k[idx_key])] = _dev(np.array(k[idx_key], np.uint32))
            li = torch.tensor([len(k[idx_key])], dtype=torch.int32, device="cuda")
            ni, ne, nl = (torch.zeros(8, dtype=torch.int32, device="cuda") for _ in range(3))
            # the lists of test.cpp:164-184 request 5 entries: a duplicated hidden entry makes list B grow from 3 to 4
            assert lib.apbf_apply_hidden_edit(h, edit.data_ptr(), le.data_ptr(), len(k["edit"]), idx.data_ptr(), li.data_ptr(), 5, 5,
                                              ni.data_ptr(), ne.data_ptr(), nl.data_ptr()) == 0
            assert u32(ni)[: int(nl[0].item())].tolist() == k[exp_key], (name, idx_key)
    k = KATS["zcurve_cells_6bit_3d"]
    pos = np.zeros((len(k["cells"]), 4), np.int32)
    pos[:, :3] = ((np.array(k["cells"], np.float32) + 0.5) * 262144.0).astype(np.int32)
    out = torch.zeros(len(k["cells"]), dtype=torch.int32, device="cuda")
    ln = torch.tensor([len(k["cells"])], dtype=torch.int32, device="cuda")
    import ctypes as C
    f3 = lambda v: (C.c_float * 3)(*v)
    assert lib.apbf_calculate_position_hash(h, _dev(pos).data_ptr(), out.data_ptr(), ln.data_ptr(), len(k["cells"]), f3((0, 0, 0)), f3((64, 64, 64)), 6) == 0
    assert u32(out).tolist() == k["expected"]
    k = KATS["position_code"]
    pos = np.array([k["position"] + [0]], np.int32)
    out = torch.zeros(1, dtype=torch.int32, device="cuda")
    one = torch.tensor([1], dtype=torch.int32, device="cuda")
    assert lib.apbf_calculate_position_code(h, _dev(np.array([0], np.uint32)).data_ptr(), _dev(pos).data_ptr(), out.data_ptr(), one.data_ptr(), 1, 0) == 0
    assert int(u32(out)[0]) == k["section0"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(INDEX["scenes"]))
def test_cuda_matches_frozen_outputs(gpu, orc, name):
    meta = INDEX["scenes"][name]
    sc = _scene(name)
    g = np.load(os.path.join(HERE, name + ".npz"))
    ctx = gpu.Context(dims=sc.dims)
    ctx.set_settings(gpu.Settings.from_buffer_copy(bytes(_settings(orc, meta, sc))))
    cap = sc.n * (700 if meta["adaptive"] else 80)
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=cap)
    aux = gpu.neighborhood_green(ctx).set_data(L).set_range_scale(meta["scale"]).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply(debug=True)
    assert np.array_equal(aux["sorted_key"], g["sorted_hash"]) and np.array_equal(aux["sorted_index"], g["sorted_index"])
    assert np.array_equal(aux["cell_start"], g["cell_start"]) and np.array_equal(aux["cell_end"], g["cell_end"])
    pairs = L.read_pairs()
    assert np.array_equal(pairs, g["pairs"])                                     # bit-exact, same grouped discovery order
    off = aux["pair_offsets"].astype(np.int64)
    assert off[-1] == len(pairs) and np.array_equal(np.diff(off), np.bincount(g["pairs"][:, 0], minlength=sc.n))
    assert np.array_equal(L.read("position"), g["position_sorted"])
    if meta["adaptive"]:
        kwfx = gpu.spread_kernel_width(ctx).set_data(L).apply(debug=True)
        assert np.array_equal(kwfx, g["kw_fixed"]) and np.array_equal(L.read_pairs(), g["kept_pairs"])
        assert np.array_equal(L.read("kernel_width"), g["kernel_width"])
    # update_transfers on the searched state: integer minima and a few unfused float operations -> bit exact.  It rewrites
    # boundary distance, target radius and boundariness only; none of them enters the quantities compared below.
    nearest = gpu.update_transfers(ctx).set_data(L).apply(debug=True)
    assert np.array_equal(nearest, g["ut_nearest"])
    for k in ("boundary_distance", "target_radius", "boundariness"):
        assert np.array_equal(L.read(k), g["ut_" + k]), k
    kw_before = L.read("kernel_width").copy()
    gpu.kernel_width_from_boundary_distance(ctx, L)
    assert np.array_equal(L.read("kernel_width"), g["ut_kernel_width"])
    L.write("kernel_width", kw_before)                # the solver below runs with the widths the fixture was made with
    a = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
    # integer accumulators in units of 2^-18: CUDA's expf/powf differ from glibc's in the last bit, which the float -> fixed
    # truncation turns into at most one unit per pair; tolerance 1 unit + 1e-5 relative (north_star)
    for k in ("density", "sq_grad_sum"):
        d = np.abs(a[k].astype(np.int64) - g[k].astype(np.int64))
        assert np.all(d <= 1 + 1e-5 * g[k].astype(np.float64)), (k, int(d.max()))
    d = np.abs(a["grad_sum"].astype(np.int64) - g["grad_sum"].astype(np.int64))
    assert np.all(d <= 3 + 1e-5 * np.abs(g["grad_sum"]).max()), int(d.max())  # a handful of the ~100 pairs of a particle may each differ by one unit
    shift = g["position_after"][:, :3].astype(np.int64) - g["position_sorted"][:, :3]
    err = np.abs(L.read("position")[:, :3].astype(np.int64) - g["position_after"][:, :3])
    assert err.max() <= 8 + 1e-5 * np.abs(shift).max(), (int(err.max()), int(np.abs(shift).max()))
