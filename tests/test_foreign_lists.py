"""GPU tests: the operators over a pair list take ANY pbd::neighbors list (source/incompressibility.h:11,
spread_kernel_width.h:10, update_transfers.h) -- not only the one this context's search produced.

incompressibility_1.comp:38-45, kernel_width.comp:27-33 and find_split_and_merge_1.comp:18-35 run one invocation per pair and
read the ids straight from the buffer: the order of the list, who wrote it, and whether a pair's mirror image is in it do not
matter to the reference.  Here the sweeps run on a grouped form of the list; for a list the context has not seen (or one that
was edited) that form is rebuilt on the device from the (id, idN) pairs (csrc/nbrlist.cu).  Each case feeds the SAME list to
the oracle's per-pair loops and to the CUDA operators.
"""
import ctypes as C

import numpy as np
import pytest

from apbf_b200 import scenes
from conftest import oracle_state
from test_gpu_parity import _check_incompressibility

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import apbf_b200
    return apbf_b200


def _scene():
    return scenes.uniform_block(18, jitter=0.25, shuffle=False)


def _variants(orc, sc):
    """pair lists over the UNSORTED scene state that no search of the context under test has produced"""
    a = sc.arrays
    brute = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 1.0, sc.n * 80)   # (id asc, idN asc)
    rng = np.random.default_rng(11)
    shuffled = brute[rng.permutation(len(brute))]
    edited = shuffled.copy()
    edited = edited[rng.random(len(edited)) > 0.2]            # every fifth pair gone: many pairs lose their mirror image
    edited = np.concatenate([edited, edited[:500]])           # 500 pairs twice (the reference would process them twice)
    edited = np.concatenate([edited, np.array([[3, sc.n + 5], [sc.n + 9, 2]], np.uint32)])   # ids beyond the list: dropped
    return dict(brute=brute, shuffled=shuffled, edited=edited)


@pytest.mark.parametrize("which", ["brute", "shuffled", "edited"])
def test_incompressibility_on_a_foreign_list(gpu, orc, which):
    sc = _scene()
    pairs = _variants(orc, sc)[which]
    s = orc.default_settings()
    st = oracle_state(orc, sc)
    before = st.position.copy()
    inside = pairs[(pairs[:, 0] < sc.n) & (pairs[:, 1] < sc.n)]
    ea = orc.incompressibility_apply(st, s, 3, inside, want_aux=True)
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=len(pairs) + 100)
    L.write_pairs(pairs)
    ga = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
    _check_incompressibility(ga, ea, L.read("position"), st.position, before)
    assert np.array_equal(L.read_pairs(), pairs)              # the caller's list is left as it is


def test_spread_kernel_width_on_a_shuffled_list(gpu, orc):
    sc = scenes.waterdrop(12, jitter=0.1)
    a = sc.arrays
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0
    pairs = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 1.5, sc.n * 700)
    pairs = pairs[np.random.default_rng(3).permutation(len(pairs))]
    st = oracle_state(orc, sc)
    ekept, ekw = orc.spread_kernel_width_apply(st, s, pairs)
    ctx = gpu.Context()
    ctx.set_settings(gpu.Settings.from_buffer_copy(bytes(s)))
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=len(pairs) + 10)
    L.write_pairs(pairs)
    gkw = gpu.spread_kernel_width(ctx).set_data(L).apply(debug=True)
    assert np.array_equal(gkw, ekw)                                            # atomicMax targets: exact
    got = L.read_pairs()
    key = lambda p: np.sort(p[:, 0].astype(np.uint64) << np.uint64(32) | p[:, 1].astype(np.uint64))
    assert np.array_equal(key(got), key(ekept))                                # kept pairs (the reference's order is arbitrary)
    assert np.array_equal(L.read("kernel_width"), st.kernel_width)
    before = st.position.copy()                                                # ... and the solver runs on the pruned list
    ea = orc.incompressibility_apply(st, s, 3, ekept, want_aux=True)
    ga = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
    _check_incompressibility(ga, ea, L.read("position"), st.position, before)


def test_update_transfers_on_a_shuffled_list(gpu, orc):
    sc = _scene()
    a = sc.arrays
    s = orc.default_settings()
    pairs = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 1.0, sc.n * 80)
    pairs = pairs[np.random.default_rng(5).permutation(len(pairs))]
    st = oracle_state(orc, sc)
    orc.update_transfers_apply(st, s, pairs)
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=len(pairs) + 10)
    L.write_pairs(pairs)
    gpu.update_transfers(ctx).set_data(L).apply()
    for k in ("boundary_distance", "target_radius", "boundariness"):
        assert np.array_equal(L.read(k), getattr(st, k)), k


def test_in_place_edit_after_a_search_is_seen(gpu, orc):
    """search -> incompressibility -> the caller rewrites the pair buffer in place -> incompressibility works on the NEW list"""
    sc = scenes.uniform_block(16, jitter=0.2, shuffle=True)
    s = orc.default_settings()
    st = oracle_state(orc, sc)
    epairs = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 80)
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=sc.n * 80)
    gpu.neighborhood_green(ctx).set_data(L).set_range_scale(1.0).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply()
    inc = gpu.incompressibility(ctx).set_data(L)
    orc.incompressibility_apply(st, s, 3, epairs)
    inc.apply()
    assert np.abs(L.read("position").astype(np.int64) - st.position).max() <= 8
    edited = epairs[::2][::-1].copy()                     # half the pairs, back to front
    st2 = orc.State(**L.read_all())                       # both arms continue from the GPU state
    before = st2.position.copy()
    ea = orc.incompressibility_apply(st2, s, 3, edited, want_aux=True)
    L.write_pairs(edited)
    ga = inc.apply(debug=True)
    _check_incompressibility(ga, ea, L.read("position"), st2.position, before)


def test_two_lists_in_one_context(gpu, orc):
    """one context, one fluid, two pair lists used in turn: each call sees the list it is given"""
    sc = _scene()
    a = sc.arrays
    s = orc.default_settings()
    near = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 0.8, sc.n * 80)
    far = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 1.0, sc.n * 80)
    assert len(near) < len(far)
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=16)
    lists = {"near": gpu.NeighborList(ctx, len(far) + 10, near), "far": gpu.NeighborList(ctx, len(far) + 10, far)}
    st = oracle_state(orc, sc)
    for which in ("near", "far", "near", "far"):
        pairs = near if which == "near" else far
        before = st.position.copy()
        ea = orc.incompressibility_apply(st, s, 3, pairs, want_aux=True)
        L.use_neighbors(lists[which])
        ga = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
        for k in ("density", "sq_grad_sum"):
            assert np.abs(ga[k].astype(np.int64) - ea[k].astype(np.int64)).max() <= 64, (which, k)
        st = orc.State(**L.read_all())                    # keep both arms on the same state
        assert np.abs(st.position[:, :3].astype(np.int64) - before[:, :3]).max() > 0
    # the lists themselves were never touched
    assert np.array_equal(lists["near"].read(), near) and np.array_equal(lists["far"].read(), far)


def test_list_written_through_the_library_is_seen(gpu, orc):
    """apbf_copy_bytes into a pair buffer (what gpu_list::operator= / apply_edit end up calling) invalidates what is remembered"""
    import torch
    sc = _scene()
    a = sc.arrays
    s = orc.default_settings()
    full = orc.brute_force_pairs(a["index_list"], a["position"], a["kernel_width"], 1.0, sc.n * 80)
    part = full[: len(full) // 3].copy()
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, neighbor_capacity=len(full) + 10)
    L.write_pairs(full)
    gpu.incompressibility(ctx).set_data(L).apply()
    src = torch.from_numpy(part.view(np.int32)).cuda()
    nb = L.neighbors()
    assert ctx.lib.apbf_copy_bytes(ctx.handle, src.data_ptr(), nb.pairs, part.nbytes) == 0
    L.words[2] = len(part)
    st = orc.State(**L.read_all())
    before = st.position.copy()
    ea = orc.incompressibility_apply(st, s, 3, part, want_aux=True)
    ga = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
    _check_incompressibility(ga, ea, L.read("position"), st.position, before)


def test_sim_hands_out_the_pair_list_on_demand(gpu, orc):
    """apbf_sim_substep keeps the grouped structure only; apbf_sim_neighbors() writes the (id, idN) list when asked"""
    sc = scenes.uniform_block(16, jitter=0.2, shuffle=True, wall_gap=3.0)
    s = orc.default_settings()
    st = oracle_state(orc, sc)
    epairs = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 64)
    ctx = gpu.Context()
    sim = gpu.Sim(ctx, sc, neighbor_capacity=sc.n * 64, solver_iterations=0)
    sim.upload(sc.arrays)
    sim.substep(1)
    assert np.array_equal(sim.read_pairs(), epairs)
    sim.substep(1)                                        # and again after the next search (the list is rewritten on demand)
    assert len(sim.read_pairs()) == sim.neighbor_count()


def test_list_grows_after_the_search(gpu, orc):
    """particles appended behind the list's length after a search (what update_transfers' splits do) have no pairs"""
    sc = scenes.uniform_block(12, jitter=0.2, shuffle=True)
    s = orc.default_settings()
    st = oracle_state(orc, sc)
    epairs = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 80)
    ctx = gpu.Context()
    L = gpu.ParticleLists(ctx, sc.arrays, capacity=sc.n + 64, neighbor_capacity=sc.n * 80)
    gpu.neighborhood_green(ctx).set_data(L).set_range_scale(1.0).set_position_range(sc.min_pos, sc.max_pos, sc.res_log2).apply()
    got = L.read_all()
    extra = 40                                            # copies of the first particles, far away from everybody
    grown = {k: np.concatenate([v, v[:extra]]) for k, v in got.items()}
    grown["index_list"] = np.arange(sc.n + extra, dtype=np.uint32)
    grown["position"][sc.n:, :3] += 262144 * 1000
    for k, v in grown.items():
        L.write(k, v)
    L.words[0] = sc.n + extra
    L.words[1] = sc.n + extra
    st2 = orc.State(**grown)
    ea = orc.incompressibility_apply(st2, s, 3, epairs, want_aux=True)
    ga = gpu.incompressibility(ctx).set_data(L).apply(debug=True)
    assert len(ga["density"]) == sc.n + extra
    assert np.abs(ga["density"].astype(np.int64) - ea["density"].astype(np.int64)).max() <= 2
    assert np.array_equal(ga["density"][sc.n:], ea["density"][sc.n:])      # self contribution only
