"""CPU tests: the oracle against the reference's own known-answer vectors and cross-checks (no GPU).

KAT sources (paths relative to the reference tree): source/test.cpp:91-105 (apply_edit), :186-197 / :269-282
(prefix sum), :199-267 (long prefix sums, property), :284-305 / :402-425 (sort), :343-364 (sort small values),
:307-341 / :366-400 (sort many values, property = std::stable_sort), :560-563 (three-particle neighbour set-up of the
disabled search tests), :623 (Z-curve cell codes of the dead sortByPositions test).
"""
import numpy as np
import pytest

from apbf_b200 import scenes
from conftest import oracle_state


def test_apply_edit_kat(orc):  # test.cpp:91-105
    out = orc.apply_edit(np.array([77, 3, 9999, 4294967295, 0], np.uint32), [1, 3, 1, 4])
    assert out.tolist() == [3, 4294967295, 3, 0]


def test_prefix_sum_kat(orc):  # test.cpp:186-197, 269-282 (inclusive)
    out = orc.prefix_sum([43, 1, 4567, 0, 1, 0, 84523487])
    assert out.tolist() == [43, 44, 4611, 4611, 4612, 4612, 84528099]


@pytest.mark.parametrize("n,mask", [(1000, 3), (512 * 512 + 1000, 1)])  # test.cpp:199-267
def test_long_prefix_sum(orc, n, mask):
    v = (np.random.default_rng(0).integers(0, 1 << 31, n, dtype=np.uint32) & mask).astype(np.uint32)
    assert np.array_equal(orc.prefix_sum(v), np.cumsum(v, dtype=np.uint64).astype(np.uint32))


def test_prefix_sum_helper_layout(orc):  # test.cpp:225-267: helper = group sums per level + 10 words
    n = 512 * 512 + 1000
    assert orc.prefix_sum_helper_length(n) == 514 + 2 + 1 + 10
    assert orc.prefix_sum_helper_length(7) == 1 + 10
    assert orc.sort_helper_length(8) == 16 + orc.prefix_sum_helper_length(16)


def test_sort_kat(orc):  # test.cpp:284-305, 402-425: stability pinned by the two 2s
    k, v = orc.sort([15, 2, 1234, 2, 0, 4294967295, 1, 4294967294], range(8))
    assert k.tolist() == [0, 1, 2, 2, 15, 1234, 4294967294, 4294967295]
    assert v.tolist() == [4, 6, 1, 3, 0, 2, 7, 5]


def test_sort_small_values_kat(orc):  # test.cpp:343-364
    k, v = orc.sort([15, 2, 3, 2, 0, 14, 1, 14], range(8))
    assert k.tolist() == [0, 1, 2, 2, 3, 14, 14, 15]
    assert v.tolist() == [4, 6, 1, 3, 2, 5, 7, 0]


@pytest.mark.parametrize("n,mask", [(512 * 512 + 123, 0xFFFFFFFF), (512 * 512 + 1000, 15)])  # test.cpp:307-341, 366-400
def test_sort_many_values(orc, n, mask):
    keys = (np.random.default_rng(0).integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & np.uint32(mask))
    k, v = orc.sort(keys, np.arange(n, dtype=np.uint32))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v, order.astype(np.uint32))
    assert np.array_equal(k, keys[order])


def test_sort_upper_bound_limits_passes(orc):  # algorithms.cpp:73: only the digits below the bound are sorted
    keys = np.array([0x15, 0x21, 0x13, 0x02], np.uint32)
    k, v = orc.sort(keys, np.arange(4, dtype=np.uint32), upper_bound=15)
    assert v.tolist() == [1, 3, 2, 0]  # ordered by the low 4 bits only, stable


def test_zcurve_vectors(orc):  # test.cpp:623: cells (2,5,1),(7,0,0),(0,1,0),(63,0,62) -> 142,73,2,187241 (6 bits per axis)
    cells = np.array([[2, 5, 1], [7, 0, 0], [0, 1, 0], [63, 0, 62]], np.float32)
    # one cell per unit: min 0, max 64, res 6 -> cell index = floor(pos)
    pos = np.zeros((4, 4), np.int32)
    pos[:, :3] = ((cells + 0.5) * 262144.0).astype(np.int32)
    h = orc.position_hash(pos, (0, 0, 0), (64, 64, 64), 6, 3)
    assert h.tolist() == [142, 73, 2, 187241]


def test_position_code_vector(orc):  # calculate_position_code.comp:28-30: raw (1,2,4) -> section 0 = 1 + 16 + 256
    pos = np.array([[1, 2, 4, 0], [-1, -1, -1, 0], [1 << 21, 0, 0, 0], [0, 1 << 21, 0, 0]], np.int32)
    idx = np.arange(4, dtype=np.uint32)
    assert orc.position_code(idx, pos, 0).tolist() == [273, 0xFFFFFFFF, 0, 0]
    assert orc.position_code(idx, pos, 1).tolist() == [0, 0xFFFFFFFF, 1 << 31, 0]
    assert orc.position_code(idx, pos, 2).tolist() == [0, 0xFFFFFFFF, 0, 1]


def test_find_value_ranges(orc):
    s, e = orc.find_value_ranges(np.arange(7), [1, 1, 4, 4, 4, 6, 6], 8)
    assert s.tolist() == [0, 0, 0, 0, 2, 0, 5, 0]
    assert e.tolist() == [0, 2, 0, 0, 5, 0, 7, 0]


def _three_particles():  # test.cpp:560-563
    R = 262144
    return dict(index_list=np.arange(3, dtype=np.uint32), position=np.array([[0, 0, 0, 1], [R, 0, 0, 1], [0, R, 0, 1]], np.int32),
                velocity=np.zeros((3, 4), np.float32), inverse_mass=np.full(3, 0.125, np.float32), radius=np.ones(3, np.float32),
                pos_backup=np.zeros((3, 4), np.int32), transferring=np.zeros(3, np.uint32), target_radius=np.ones(3, np.float32),
                kernel_width=np.array([1.0, 1.0, 2.0], np.float32), boundariness=np.ones(3, np.float32),
                boundary_distance=np.zeros(3, np.uint32))


def _pairs_by_position(st, pairs):
    """pair list as a set of (position of id, position of idN) -- independent of the re-ordering"""
    pos = st.position[st.index_list][:, :3]
    return sorted((tuple(pos[a]), tuple(pos[b])) for a, b in pairs)


def test_three_particle_neighbours(orc):
    """ranges 1,1,2 -> {(0,1),(0,2),(1,0),(2,0),(2,1)}: particle 2 reaches particle 1 at distance sqrt(2), not vice versa"""
    s = orc.default_settings()
    R = 262144
    P = [(0, 0, 0), (R, 0, 0), (0, R, 0)]
    expected = sorted((P[a], P[b]) for a, b in [(0, 1), (0, 2), (1, 0), (2, 0), (2, 1)])
    st = orc.State(**_three_particles())
    bf = orc.brute_force_pairs(st.index_list, st.position, st.kernel_width, 1.0, 16)
    assert _pairs_by_position(st, bf) == expected
    st = orc.State(**_three_particles())
    g = orc.green_apply(st, s, 3, 1.0, (-10, -10, -10), (10, 10, 10), 6, 16)  # test.cpp:572
    assert _pairs_by_position(st, g) == expected
    st = orc.State(**_three_particles())
    b = orc.binary_search_apply(st, s, 1.0, 16)
    assert _pairs_by_position(st, b) == expected


def test_update_transfers_hand_derived(orc):
    """find_split_and_merge_1/2/3.comp by hand on the three particles of test.cpp:560-563 with the pair set
    {(0,1),(0,2),(1,0),(2,0),(2,1)}: unit distance = 262144 fixed-point units, sqrt(2) -> uint(370727.6) = 370727."""
    R, S2 = 262144, 370727
    st = orc.State(**_three_particles())
    st.boundary_distance[:] = [1000, 2000, 3000]
    st.boundariness[:] = [1.0, 0.25, 0.0]
    st.radius[:] = [1.0, 1.0, 0.5]
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0
    s.mUpdateTargetRadius = 1
    s.mSmallestTargetRadius, s.mTargetRadiusOffset, s.mTargetRadiusScaleFactor = 1.0, 0.5, 2.0
    pairs = np.array([[0, 1], [0, 2], [1, 0], [2, 0], [2, 1]], np.uint32)
    nearest = orc.update_transfers_apply(st, s, pairs)
    # pass 1: bd[n0] = min over pairs of old_bd[n1] + dist
    bd = [min(2000 + R, 3000 + R), 1000 + R, min(1000 + R, 2000 + S2)]
    # pass 2: nearest neighbour = the pair at the minimum distance; particle 0 has two at distance R -> the last one of the list
    assert nearest.tolist() == [2, 0, 0]
    # pass 3: target radius = smallest + max(0, (bd / 2^18 - offset) * factor); decay mix(bd, radius, boundariness >= 1)
    f32 = np.float32
    exp_tr = [f32(1.0) + max(f32(0.0), (f32(b) / f32(R) - f32(0.5)) * f32(2.0)) for b in bd]
    assert np.array_equal(st.target_radius, np.array(exp_tr, f32))
    assert st.boundariness.tolist() == [1.0, 0.0, 0.0]
    assert st.boundary_distance.tolist() == [R, bd[1], bd[2]]          # particle 0 is on the boundary: its distance falls back to its radius
    # pool.cpp:77-80 with factor 2 / 2^18, lower bound 4 * radius, step 1 %: particles 0 and 1 (width 1, value ~2) step to 1.01 and
    # are lifted to the lower bound 4; particle 2 (width 2, value 2.0076 within one step) reaches the value, above its bound 2
    orc.kernel_width_from_boundary_distance(st, s)
    assert np.array_equal(st.kernel_width, np.array([4.0, 4.0, f32(bd[2]) * (f32(2.0) / f32(R))], f32))
    st.radius[:] = 0.1
    st.kernel_width[:] = [1.0, 3.0, 1.0]
    orc.kernel_width_from_boundary_distance(st, s)
    v = [f32(b) * (f32(2.0) / f32(R)) for b in st.boundary_distance]      # 2, ~2.0076, ~2.0076
    exp = [f32(1.0) * (f32(1.0) + s.mKernelWidthAdaptionSpeed), f32(3.0) * (f32(1.0) - f32(s.mKernelWidthAdaptionSpeed)),
           f32(1.0) * (f32(1.0) + s.mKernelWidthAdaptionSpeed)]
    assert np.array_equal(st.kernel_width, np.array(exp, f32)) and all(e < x for e, x in zip(exp[::2], v[::2]))


@pytest.mark.parametrize("dims", [2, 3])
def test_searches_agree_with_brute_force(orc, dims):
    sc = scenes.uniform_block(10 if dims == 3 else 40, jitter=0.3, dims=dims, shuffle=True)
    s = orc.default_settings()
    cap = sc.n * 700
    st = oracle_state(orc, sc)
    # variable ranges: unmirrored pairs must show up
    st.kernel_width *= np.random.default_rng(3).uniform(0.6, 1.4, sc.n).astype(np.float32)
    st2, st3 = st.copy(), st.copy()
    g = orc.green_apply(st, s, dims, 1.5, sc.min_pos, sc.max_pos, sc.res_log2, cap)
    bf = orc.brute_force_pairs(st.index_list, st.position, st.kernel_width, 1.5, cap)
    assert len(g) == len(bf) and len(np.unique(g, axis=0)) == len(g)
    assert np.array_equal(np.unique(g, axis=0), np.unique(bf, axis=0))
    assert np.all(np.diff(g[:, 0].astype(np.int64)) >= 0)  # grouped by id
    if dims == 3:
        b = orc.binary_search_apply(st2, s, 1.5, cap)
        assert _pairs_by_position(st2, b) == _pairs_by_position(st, g)
    # the re-ordering is a permutation that keeps every particle's attributes together
    key = lambda t: np.lexsort(t.position[:, :3].T[::-1])
    a, c = key(st3), key(st)
    for name in ("position", "inverse_mass", "radius"):
        assert np.array_equal(getattr(st3, name)[a], getattr(st, name)[c])
    assert np.array_equal(st3.kernel_width[a], st.kernel_width[c])  # per-id arrays follow (index list is the identity)
    assert np.array_equal(st.index_list, np.arange(sc.n))


def test_green_sorted_hash_and_cells(orc):
    sc = scenes.uniform_block(10, jitter=0.2, shuffle=True)
    st = oracle_state(orc, sc)
    pairs, aux = orc.green_apply(st, orc.default_settings(), 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 64, want_aux=True)
    assert np.all(np.diff(aux["sorted_hash"].astype(np.int64)) >= 0)
    h = orc.position_hash(st.position, sc.min_pos, sc.max_pos, sc.res_log2, 3)
    assert np.array_equal(h, aux["sorted_hash"])
    occ = aux["cell_end"].astype(np.int64) - aux["cell_start"]
    assert occ.sum() == sc.n and occ.min() >= 0


def test_kernel_functions_basic(orc):
    s = orc.default_settings()
    for hk in range(5):
        s.mHeightKernelId = hk
        w0 = orc.kernel_height(s, 3, [0, 0, 0], 4.0)
        w1 = orc.kernel_height(s, 3, [1.0, 1.0, 1.0], 4.0)
        assert w0 > w1 >= 0.0
    for gk in range(5):
        s.mGradientKernelId = gk
        g = orc.kernel_gradient(s, 3, [1.0, 0.5, -0.25], 4.0)
        assert np.dot(g, [1.0, 0.5, -0.25]) < 0  # points towards the centre
        assert np.all(orc.kernel_gradient(s, 3, [0, 0, 0], 4.0) == 0)


def test_incompressibility_pushes_compressed_particles_apart(orc):
    sc = scenes.uniform_block(8, jitter=0.0)
    st = oracle_state(orc, sc)
    st.position[:, :3] = (st.position[:, :3].astype(np.float64) * 0.8).astype(np.int32)  # 20 % compression
    s = orc.default_settings()
    pairs = orc.green_apply(st, s, 3, 1.0, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 128)
    before = st.position.copy()
    aux = orc.incompressibility_apply(st, s, 3, pairs, want_aux=True)
    centre = np.argmin(np.abs(before[:, :3]).sum(1))
    assert aux["lam"][centre] < 0
    r0 = np.linalg.norm(before[:, :3].astype(np.float64), axis=1)
    r1 = np.linalg.norm(st.position[:, :3].astype(np.float64), axis=1)
    assert (r1 - r0)[r0 > 0].mean() > 0  # the block expands
    assert np.array_equal(st.position[:, 3], before[:, 3])


def test_spread_kernel_width_prunes_and_spreads(orc):
    sc = scenes.waterdrop(10, jitter=0.05)
    st = oracle_state(orc, sc)
    s = orc.default_settings()
    s.mBaseKernelWidthOnBoundaryDistance = 0
    pairs = orc.green_apply(st, s, 3, 1.5, sc.min_pos, sc.max_pos, sc.res_log2, sc.n * 400)
    old_kw = st.kernel_width.copy()
    kept, kwfx = orc.spread_kernel_width_apply(st, s, pairs)
    assert 0 < len(kept) < len(pairs)
    assert np.all(kwfx >= np.trunc(np.maximum(st.radius, st.target_radius) * 4 * 262144).astype(np.uint32))
    assert np.all(np.abs(st.kernel_width / old_kw - 1) <= s.mKernelWidthAdaptionSpeed * 1.0001)


def test_box_collision_keeps_particles_outside_walls(orc):
    sc = scenes.uniform_block(8)
    st = oracle_state(orc, sc)
    st.position[:, 0] -= 3 * 262144 // 2  # push the outermost layer 1.5 units into the (radius-inflated) -x wall
    orc.box_collision(st, sc.box_min, sc.box_max)
    x = st.position[:, 0] / 262144.0
    assert x.min() >= sc.box_max[0, 0] + 1.0 - 1e-3  # back at wall face + radius (the hash jitter only widens the wall)
    assert x.min() <= sc.box_max[0, 0] + 1.0 + 0.05 + 1e-3


def test_substep_runs_and_conserves_particles(orc):
    sc = scenes.uniform_block(8, jitter=0.2, shuffle=True)
    st = oracle_state(orc, sc)
    s = orc.default_settings()
    pairs = orc.substep(st, s, dims=3, basic_pbf=True, solver_iterations=4, min_pos=sc.min_pos, max_pos=sc.max_pos,
                        res_log2=sc.res_log2, box_min4=sc.box_min, box_max4=sc.box_max, cap=sc.n * 64)
    assert len(pairs) > 0 and st.n == sc.n
    assert sorted(st.inverse_mass.tolist()) == sorted(sc.arrays["inverse_mass"].tolist())


def test_incompressibility_two_particles_closed_form(orc):
    """Two equal particles 1.5 apart, kernel width 3, Gauss kernels, D = 3 -- every quantity of incompressibility_0..3.comp in
    closed form (float64), independent of the restatement's code:
      W(r) = H exp(-r^2 pi H^(2/3)), H = 0.6 / (h/2)^3                              kernels.glsl:84-89,104
      density = m (W(0) + W(d)); g = -2 W(d) d pi H^(2/3) along the pair           incompressibility_0.comp:43, _1.comp:52-58, kernels.glsl:91-97
      lambda = (1 - density) / (|g|^2 m + |m g|^2 / m + 0.01)   (rest density 1)   incompressibility_2.comp:72-98
      own shift = m g lambda / m, neighbour shift = g lambda                       incompressibility_2.comp:106-109, _3.comp:66
    Both particles end up |g lambda| further out on each side, twice (own shift + the neighbour's push)."""
    d, h, m = 1.5, 3.0, 8.0
    pos = np.zeros((2, 4), np.int32)
    pos[1, 0] = int(d * 262144)
    st = orc.State(index_list=np.arange(2, dtype=np.uint32), position=pos.copy(), velocity=np.zeros((2, 4), np.float32),
                   inverse_mass=np.full(2, 1.0 / m, np.float32), radius=np.ones(2, np.float32), pos_backup=pos.copy(),
                   transferring=np.zeros(2, np.uint32), target_radius=np.ones(2, np.float32), kernel_width=np.full(2, h, np.float32),
                   boundariness=np.ones(2, np.float32), boundary_distance=np.full(2, 262144, np.uint32))
    s = orc.default_settings()
    s.mHeightKernelId, s.mGradientKernelId, s.mSmallestTargetRadius = 1, 1, 1.0
    aux = orc.incompressibility_apply(st, s, 3, np.array([[0, 1], [1, 0]], np.uint32), want_aux=True)
    H = 0.6 / (h / 2.0) ** 3
    iv = np.pi * H ** (2.0 / 3.0)
    W0, Wd = H, H * np.exp(-d * d * iv)
    g = 2.0 * Wd * d * iv                                             # magnitude; points from the neighbour towards the particle
    density = m * (W0 + Wd)
    lam = (1.0 - density) / (g * g * m + (m * g) ** 2 / m + 0.01)
    assert density > 1.0 and lam < 0.0
    assert np.all(np.abs(aux["density"].astype(np.int64) - int(density * 262144)) <= 2)
    assert np.all(np.abs(np.abs(aux["grad_sum"][:, 0].astype(np.int64)) - int(m * g * 262144)) <= 2) and np.all(aux["grad_sum"][:, 1:] == 0)
    assert aux["grad_sum"][0, 0] < 0 < aux["grad_sum"][1, 0]            # particle 0 sees its neighbour at +x: gradient along -x
    assert np.all(np.abs(aux["sq_grad_sum"].astype(np.int64) - int(g * g * m * 262144)) <= 2)
    # lambda is formed from the truncated fixed-point sums (incompressibility_2.comp:67-69): the closed form to 1e-3, the same
    # expression over the integers checked above to float accuracy
    assert np.allclose(aux["lam"], lam, rtol=1e-3)
    dq, gq, sq = aux["density"][0] / 262144.0, aux["grad_sum"][0, 0] / 262144.0, aux["sq_grad_sum"][0] / 262144.0
    lam = (1.0 - dq) / (sq + gq * gq / m + 0.01)
    assert np.allclose(aux["lam"], lam, rtol=1e-5)
    shift = 2.0 * g * (-lam) * 262144                                 # own shift + the push from the neighbour's pair
    assert abs((pos[0, 0] - st.position[0, 0]) - shift) <= 3 and abs((st.position[1, 0] - pos[1, 0]) - shift) <= 3
    assert np.all(st.position[:, 1:3] == 0)


def test_box_collision_exact_push_out(orc):
    """id 0 at (1.2, 0, 0) with radius 1, box [1, 9] x [-5, 5]^2 (box_collision.comp:44-57): the jitter of the min faces is
    hash31(0 * x - 0 - 0) = hash31(0) = 0, so the inflated min face sits at exactly 0; penetration 1.2 along x against about 6
    along y and z and 8.8 to the max face -> the particle lands on x = 0 exactly.  The same particle outside the inflated box
    (x = -0.5) only goes through the re-quantisation (:59)."""
    def run(x):
        pos = np.zeros((1, 4), np.int32)
        pos[0, 0] = int(round(x * 262144))
        st = orc.State(index_list=np.zeros(1, np.uint32), position=pos, velocity=np.zeros((1, 4), np.float32), inverse_mass=np.full(1, 0.125, np.float32),
                       radius=np.ones(1, np.float32), pos_backup=pos.copy(), transferring=np.zeros(1, np.uint32), target_radius=np.ones(1, np.float32),
                       kernel_width=np.full(1, 4.0, np.float32), boundariness=np.ones(1, np.float32), boundary_distance=np.full(1, 262144, np.uint32))
        orc.box_collision(st, np.array([[1, -5, -5, 0]], np.float32), np.array([[9, 5, 5, 0]], np.float32))
        return st.position[0, :3].tolist()
    assert run(1.2) == [0, 0, 0]
    assert run(-0.5) == [-131072, 0, 0]


def _pair_state(orc, xs, radii, widths):
    n = len(xs)
    pos = np.zeros((n, 4), np.int32)
    pos[:, 0] = [int(round(x * 262144)) for x in xs]
    radii = np.asarray(radii, np.float32)
    return orc.State(index_list=np.arange(n, dtype=np.uint32), position=pos, velocity=np.zeros((n, 4), np.float32),
                     inverse_mass=(1.0 / (2.0 * radii) ** 3).astype(np.float32), radius=radii, pos_backup=pos.copy(),
                     transferring=np.zeros(n, np.uint32), target_radius=radii.copy(), kernel_width=np.asarray(widths, np.float32),
                     boundariness=np.ones(n, np.float32), boundary_distance=(radii * 262144).astype(np.uint32))


def test_spread_kernel_width_hand_derived(orc):
    """radius 1 (width 4) and radius 2 (width 8) six units apart (kernel_width_init.comp:33-35, kernel_width.comp:46-60,
    uint_to_float_but_gradual.comp:21-38):
      pair (0,1): 6 - 4 = 2 beyond the small kernel, 2 / (4 * 0.5) = 1 -> influence 0; 6 > max(4, 4) -> the pair is dropped
      pair (1,0): inside the large kernel -> influence 1 -> particle 0 is asked for width 8; 6 <= 8 -> the pair stays
    fixed-point widths {8, 8} * 2^18; particle 0 moves from 4 towards 8 by one relative step, particle 1 stays at 8."""
    st = _pair_state(orc, [0.0, 6.0], [1.0, 2.0], [4.0, 8.0])
    s = orc.default_settings()
    s.mBaseKernelWidthOnTargetRadius = 0
    kept, kwfx = orc.spread_kernel_width_apply(st, s, np.array([[0, 1], [1, 0]], np.uint32))
    assert kept.tolist() == [[1, 0]] and kwfx.tolist() == [8 * 262144, 8 * 262144]
    step = np.float32(4.0) * (np.float32(1.0) + np.float32(s.mKernelWidthAdaptionSpeed))
    assert 4.0 < step < 8.0 and st.kernel_width.tolist() == [step, 8.0]
    # half way into the fall-off: 5 units apart -> (5 - 4) / 2 = 0.5 -> the large particle is asked for 4 * 0.5 = 2 (below its own 8)
    st = _pair_state(orc, [0.0, 5.0], [1.0, 2.0], [4.0, 8.0])
    kept, kwfx = orc.spread_kernel_width_apply(st, s, np.array([[0, 1]], np.uint32))
    assert kept.tolist() == [] and kwfx.tolist() == [4 * 262144, 8 * 262144]


def test_velocity_handling_hand_derived(orc):
    """infer_velocity.comp:32, apply_acceleration.comp:28, apply_velocity.comp:31 with mLastDeltaTime == dt (velocity_handling.cpp:
    15-31): v = (pos - backup) / (dt 2^18), v += a dt, pos += ivec3(v (dt 2^18)), backup = the old position"""
    dt = np.float32(1.0 / 60.0)
    st = _pair_state(orc, [0.5], [1.0], [4.0])
    st.position[0, 1] = 65536
    st.pos_backup[:] = 0
    orc.velocity_handling(st, float(dt), (0.0, -10.0, 0.0))
    k = dt * np.float32(262144.0)
    v = np.array([131072, 65536, 0], np.float32) / k
    v[1] += np.float32(-10.0) * dt
    assert st.velocity[0, :3].tolist() == v.tolist() and abs(float(v[0]) - 30.0) < 1e-4 and abs(float(v[1]) - (15.0 - 1.0 / 6.0)) < 1e-4
    assert st.pos_backup[0, :3].tolist() == [131072, 65536, 0]
    assert st.position[0, :3].tolist() == [131072 + int(v[0] * k), 65536 + int(v[1] * k), 0]
    assert abs(st.position[0, 0] - 262144) <= 1 and abs(st.position[0, 1] - (131072 - 728)) <= 1   # 0.5 + 30/60 ; 0.25 + 14.8333/60


@pytest.mark.parametrize("dims", [3, 2])
def test_kernels_are_normalised_and_gradients_are_their_derivatives(orc, dims):
    """kernels.glsl:3-123, independent of the restatement's code: every height kernel integrates to 1 over its support in 3-D
    (Gauss: no cut-off, integrated to 2.5 h); in 2-D the dimension-aware ones do (Gauss, cone, quadratic spike) while cubic and
    poly6 keep their 3-D constants (the reference's behaviour: 0.70 and 0.62); the gradient kernel with the same id is the
    radial derivative of the height kernel (cubic, Gauss, cone, quadratic spike; id 2 pairs poly6 with spiky)."""
    s = orc.default_settings()
    h = 2.0
    for hk in range(5):
        s.mHeightKernelId = s.mGradientKernelId = hk
        rs = np.linspace(1e-4, 2.5 * h if hk == 1 else h * 1.0001, 3000)
        w = np.array([orc.kernel_height(s, dims, [r, 0, 0], h) for r in rs])
        integral = np.trapezoid(w * (4 * np.pi * rs ** 2 if dims == 3 else 2 * np.pi * rs), rs)
        expected = 1.0 if dims == 3 or hk in (1, 3, 4) else (0.7 if hk == 0 else 0.6152)
        assert abs(integral - expected) < 2e-3, (hk, integral)
        if hk == 2:
            continue
        for r in (0.3 * h, 0.6 * h, 0.9 * h):
            fd = (orc.kernel_height(s, dims, [r + 1e-3, 0, 0], h) - orc.kernel_height(s, dims, [r - 1e-3, 0, 0], h)) / 2e-3
            g = orc.kernel_gradient(s, dims, [r, 0, 0], h)
            assert abs(g[0] - fd) <= 2e-3 * abs(fd) + 1e-5 and g[1] == 0 and g[2] == 0, (hk, r, g, fd)
