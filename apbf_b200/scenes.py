"""Synthetic particle states for the BASELINE.json configs (host-side numpy, no device work).

Seeding restates pbd::initialize::add_box_shape (source/initialize.cpp:5-33 with
shaders/initialize_box.comp:30-50): a regular lattice with spacing 2r, positions stored as
int32 fixed point = trunc(pos * 2^18), inverse mass 1/(2r)^D, radius r, velocity 0; fluid arrays as in
pool::pool (source/pool.cpp:20-25): kernel_width = 4r, target_radius = r, boundariness = 1,
boundary_distance = uint(r * 2^18).  Pool walls follow source/pool.cpp:30-40.
"""
from dataclasses import dataclass, field

import numpy as np

POS_RESOLUTION = 262144.0
KERNEL_SCALE = 4.0


@dataclass
class Scene:
    name: str
    dims: int
    arrays: dict                 # the reference's lists as numpy arrays (see oracle.State.FIELDS)
    min_pos: tuple               # Green grid bounds (neighborhood_green::set_position_range)
    max_pos: tuple
    res_log2: int
    box_min: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    box_max: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    basic_pbf: bool = True
    solver_iterations: int = 4
    smallest_target_radius: float = 1.0

    @property
    def n(self):
        return self.arrays["index_list"].shape[0]


def _lattice_state(counts, origin, r, jitter=0.0, seed=1234, dims=3, radius_of=None):
    """counts particles per axis starting at `origin` (float32 arithmetic as initialize_box.comp:43-45)."""
    cx, cy, cz = [int(c) for c in counts]
    n = cx * cy * cz
    ids = np.arange(n, dtype=np.uint32)
    gx = ids // (cz * cy)
    rem = ids - gx * (cz * cy)
    gy = rem // cz
    gz = rem - gy * cz
    g = np.stack([gx, gy, gz], axis=1).astype(np.float32)
    pos = np.asarray(origin, np.float32)[None, :] + np.float32(2.0) * np.float32(r) * g
    if jitter:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape).astype(np.float32)
    ipos = np.zeros((n, 4), np.int32)
    ipos[:, :3] = np.trunc(pos.astype(np.float32) * np.float32(POS_RESOLUTION)).astype(np.int32)
    if dims < 3:
        ipos[:, 2] = 0
    radius = np.full(n, r, np.float32) if radius_of is None else radius_of(pos).astype(np.float32)
    inv_mass = (np.float32(1.0) / np.power(np.float32(2.0) * radius, np.float32(dims))).astype(np.float32)
    return dict(
        index_list=ids.copy(),
        position=ipos,
        velocity=np.zeros((n, 4), np.float32),
        inverse_mass=inv_mass,
        radius=radius,
        pos_backup=ipos.copy(),
        transferring=np.zeros(n, np.uint32),
        target_radius=radius.copy(),
        kernel_width=(radius * np.float32(KERNEL_SCALE)).astype(np.float32),
        boundariness=np.ones(n, np.float32),
        boundary_distance=np.trunc(radius * np.float32(POS_RESOLUTION)).astype(np.uint32),
    )


def _res_for(extent, cell):
    """smallest res with 2^res cells of size >= cell covering extent... i.e. cell size extent/2^res >= cell"""
    res = 1
    while extent / (1 << (res + 1)) >= cell:
        res += 1
    return res


def _res_near(extent, h):
    """the resolution whose cell size extent / 2^res is closest to h (in ratio): cells much larger than the search range add
    candidates to every distance-test batch, cells much smaller add cells to every walk (measured: profiles/r02_variants.md)"""
    best, res = None, 1
    for k in range(1, 10):
        d = abs(np.log((extent / (1 << k)) / h))
        if best is None or d < best:
            best, res = d, k
    return res


def pool_walls(mn, mx, r, dims=3, wall=4.0):
    """source/pool.cpp:30-40"""
    mn = np.asarray(mn, np.float32); mx = np.asarray(mx, np.float32)
    w_min = mn - np.float32(wall)
    w_max = mx + np.float32(wall) + np.array([0, 2 * r, 0], np.float32)
    boxes = [
        (w_min, (w_min[0] + wall, w_max[1], w_max[2])),
        (w_min, (w_max[0], w_min[1] + wall, w_max[2])),
        ((w_max[0] - wall, w_min[1], w_min[2]), w_max),
        ((w_min[0], w_max[1] - wall, w_min[2]), w_max),
    ]
    if dims > 2:
        boxes += [(w_min, (w_max[0], w_max[1], w_min[2] + wall)), ((w_min[0], w_min[1], w_max[2] - wall), w_max)]
    bmin = np.zeros((len(boxes), 4), np.float32); bmax = np.zeros((len(boxes), 4), np.float32)
    for i, (a, b) in enumerate(boxes):
        bmin[i, :3] = a; bmax[i, :3] = b
    return bmin, bmax


def uniform_block(side=64, r=1.0, jitter=0.0, seed=1234, dims=3, res_log2=None, shuffle=False, wall_gap=0.0):
    """configs[0]/[4]: side^D lattice centred on the origin, fixed kernel width 4r, basic PBF, 4 iterations.
    Grid bounds = particle bounds +- 4r.  wall_gap moves the pool walls away from the block: the wall
    jitter of box_collision.comp:46-47 is a chaotic hash of the position, so parity runs over several iterations keep
    the particles out of wall contact (box collision itself is compared on identical inputs)."""
    counts = (side, side, side if dims == 3 else 1)
    half = side * r  # particle centres span [-(side-1)r, (side-1)r]
    origin = (-(side - 1) * r, -(side - 1) * r, -(side - 1) * r if dims == 3 else 0.0)
    arrays = _lattice_state(counts, origin, r, jitter, seed, dims)
    lo = -(half + 4 * r); hi = half + 4 * r
    if res_log2 is None:   # (blocks below 100^3 keep the rule the golden fixtures of tests/golden were seeded with)
        res_log2 = _res_near(hi - lo, 4.0 * r) if side >= 100 else _res_for(hi - lo, 4.0 * r)
    if shuffle:
        arrays = shuffle_state(arrays, seed)
    sc = Scene(name=f"uniform_{side}^{dims}" + ("_jitter" if jitter else ""), dims=dims, arrays=arrays,
               min_pos=(lo, lo, lo), max_pos=(hi, hi, hi), res_log2=res_log2, basic_pbf=True,
               solver_iterations=4, smallest_target_radius=r)
    sc.box_min, sc.box_max = pool_walls((-half - wall_gap,) * 3, (half + wall_gap,) * 3, r, dims)
    return sc


def uniform_block_brick(side, rank, world, r=1.0, jitter=0.1, seed=1234, res_log2=None):
    """The particles of uniform_block(side) that ONE rank of a brick partition owns (apbf_b200/multi_gpu.py: the bricks are the
    halves of the search grid along z, then y, then x), generated without ever holding the whole scene: at 400^3 = 64 M
    particles eight ranks each seeding the full lattice would need ~100 GB of host memory.  `side` must be even: the block is
    centred on the origin and the cuts are the coordinate planes, so a lattice site belongs to the half its index says (the
    jitter is smaller than r).  The jitter is a hash of the global lattice index: the union over the ranks is one well-defined
    scene, whatever the number of ranks.  Returns (scene with the brick's arrays, total number of particles)."""
    assert side % 2 == 0 and world in (1, 2, 4, 8) and jitter < r
    lw = {1: 0, 2: 1, 4: 2, 8: 3}[world]
    lo, hi = [0, 0, 0], [side, side, side]
    for b in range(lw):
        axis = 2 - b                                  # bit 0 of the rank (MSB first): z, then y, then x
        half = (rank >> (lw - 1 - b)) & 1
        lo[axis], hi[axis] = (side // 2, side) if half else (0, side // 2)
    gx, gy, gz = np.meshgrid(np.arange(lo[0], hi[0], dtype=np.uint32), np.arange(lo[1], hi[1], dtype=np.uint32),
                             np.arange(lo[2], hi[2], dtype=np.uint32), indexing="ij")
    g = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    del gx, gy, gz
    n = g.shape[0]
    gid = (g[:, 0].astype(np.uint64) * np.uint64(side) + g[:, 1].astype(np.uint64)) * np.uint64(side) + g[:, 2].astype(np.uint64)
    pos = np.float32(-(side - 1) * r) + np.float32(2.0 * r) * g.astype(np.float32)
    del g
    if jitter:
      with np.errstate(over="ignore"):
        for d in range(3):
            h = (gid * np.uint64(3) + np.uint64(d) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
            h ^= h >> np.uint64(33); h *= np.uint64(0xFF51AFD7ED558CCD); h ^= h >> np.uint64(33); h *= np.uint64(0xC4CEB9FE1A85EC53); h ^= h >> np.uint64(33)
            u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))          # [0, 1)
            pos[:, d] += (u * np.float32(2.0 * jitter) - np.float32(jitter)).astype(np.float32)
    del gid
    ipos = np.zeros((n, 4), np.int32)
    ipos[:, :3] = np.trunc(pos * np.float32(POS_RESOLUTION)).astype(np.int32)
    del pos
    radius = np.full(n, r, np.float32)
    arrays = dict(index_list=np.arange(n, dtype=np.uint32), position=ipos, velocity=np.zeros((n, 4), np.float32),
                  inverse_mass=np.full(n, 1.0 / (2.0 * r) ** 3, np.float32), radius=radius, pos_backup=ipos.copy(),
                  transferring=np.zeros(n, np.uint32), target_radius=radius.copy(), kernel_width=np.full(n, r * KERNEL_SCALE, np.float32),
                  boundariness=np.ones(n, np.float32), boundary_distance=np.full(n, int(r * POS_RESOLUTION), np.uint32))
    perm = np.random.default_rng(seed + rank).permutation(n)       # hidden slots in random order: exercises the sort / reorder
    arrays = {k: (v if k == "index_list" else v[perm]) for k, v in arrays.items()}
    half = side * r
    glo, ghi = -(half + 4 * r), half + 4 * r
    if res_log2 is None:
        res_log2 = _res_near(ghi - glo, 4.0 * r)
    sc = Scene(name=f"uniform_{side}^3_jitter_brick{rank}of{world}", dims=3, arrays=arrays, min_pos=(glo,) * 3, max_pos=(ghi,) * 3, res_log2=res_log2,
               basic_pbf=True, solver_iterations=4, smallest_target_radius=r)
    sc.box_min, sc.box_max = pool_walls((-half,) * 3, (half,) * 3, r, 3)
    return sc, side ** 3


def waterfall_boxes(mn, mx, r, dims=3, wall=4.0):
    """source/waterfall.cpp:28-48: the top pool (closed: left, floor, right, lid, and the two z walls in 3-D) plus the bottom pool,
    shifted by wMin - wMax in x and y (left, floor, right, two z walls) -- 11 boxes in 3-D, 7 in 2-D"""
    mn = np.asarray(mn, np.float32); mx = np.asarray(mx, np.float32)
    w_min = mn - np.float32(wall)
    w_max = mx + np.float32(wall) + np.array([0, 2 * r, 0], np.float32)
    shift = w_min - w_max
    shift[2] = 0.0
    top = [
        (w_min, (w_min[0] + wall, w_max[1], w_max[2])),
        (w_min, (w_max[0], w_min[1] + wall, w_max[2])),
        ((w_max[0] - wall, w_min[1], w_min[2]), w_max),
        ((w_min[0], w_max[1] - wall, w_min[2]), w_max),
    ]
    bottom = [(np.asarray(a, np.float32) + shift, np.asarray(b, np.float32) + shift) for a, b in top[:3]]
    boxes = top + bottom
    if dims > 2:
        z = [(w_min, (w_max[0], w_max[1], w_min[2] + wall)), ((w_min[0], w_min[1], w_max[2] - wall), w_max)]
        boxes += z + [(np.asarray(a, np.float32) + shift, np.asarray(b, np.float32) + shift) for a, b in z]
    bmin = np.zeros((len(boxes), 4), np.float32); bmax = np.zeros((len(boxes), 4), np.float32)
    for i, (a, b) in enumerate(boxes):
        bmin[i, :3] = a; bmax[i, :3] = b
    return bmin, bmax


def waterfall(nx=252, ny=252, nz=252, r=1.0, jitter=0.05, seed=17, adaptive=False, res_log2=None):
    """configs[3]: the waterfall scene (source/waterfall.cpp:6-48) -- a block of fluid filling the top pool, 11 collision boxes
    (the user opens the pool at run time in the reference; here it stays closed, what counts is the box_collision load).
    252^3 = 16 003 008 particles at full size."""
    ext = np.array([nx, ny, nz], np.float32) * 2 * r
    mn = np.zeros(3, np.float32)
    arrays = _lattice_state((nx, ny, nz), mn + r, r, jitter, seed)
    margin = 6.0 * r + 4.0
    lo = [float(v - margin) for v in mn]
    hi = [float(v + margin) for v in ext + np.array([0, 2 * r, 0], np.float32)]
    if res_log2 is None:   # cells of about one kernel width (4r): larger cells only add candidates to every distance-test batch
        res_log2 = _res_for(max(h - l for l, h in zip(lo, hi)), 3.0 * r)
    sc = Scene(name=f"waterfall_{nx}x{ny}x{nz}", dims=3, arrays=shuffle_state(arrays, seed), min_pos=tuple(lo), max_pos=tuple(hi),
               res_log2=res_log2, basic_pbf=not adaptive, solver_iterations=4, smallest_target_radius=r)
    sc.box_min, sc.box_max = waterfall_boxes(mn, ext, r, 3)
    return sc


def shuffle_state(arrays, seed=7):
    """Random permutation of the hidden slots (index list stays the identity): exercises the sort/reorder."""
    n = arrays["index_list"].shape[0]
    perm = np.random.default_rng(seed).permutation(n)
    out = {}
    for k, v in arrays.items():
        out[k] = v.copy() if k == "index_list" else v[perm].copy()
    return out


def dam_break(nx=100, ny=100, nz=100, r=1.0, jitter=0.05, seed=99, adaptive=True, blocks=1, center_y=False, res_log2=None, grid="pool"):
    """configs[1]: a block occupying one third of a pool floor (walls per pool.cpp:30-40), adaptive widths.
    Multi-GPU variants (bench.py --gpus N; bricks = halves of the grid along z, then y, then x): `blocks=2` puts a second,
    mirrored block at the far end of the pool (the pool is twice as long, the blocks still cover a third of the floor) so
    that the x halves hold the same number of particles; `center_y` makes the grid's y range symmetric about the block.
    grid: "pool" = the search grid spans the pool plus a margin per axis, like pool.cpp:48 (the golden fixtures and the parity
    scenes were seeded with it); "cube" = equal extents on all axes (what bench.py times)."""
    ext = np.array([nx, ny, nz], np.float32) * 2 * r
    pool_min = np.array([0, 0, 0], np.float32)
    pool_max = np.array([3 * blocks * ext[0], 1.5 * ext[1], ext[2]], np.float32)
    arrays = _lattice_state((nx, ny, nz), pool_min + r, r, jitter, seed)
    if blocks == 2:
        far = _lattice_state((nx, ny, nz), (pool_max[0] - ext[0] + r, r, r), r, jitter, seed + 1)
        arrays = {k: np.concatenate([arrays[k], far[k]]) for k in arrays}
        arrays["index_list"] = np.arange(arrays["position"].shape[0], dtype=np.uint32)
    margin = 6.0 * r + 4.0
    lo = [float(v - margin) for v in pool_min]
    hi = [float(v + margin) for v in pool_max]
    if center_y:
        cy = 0.5 * float(ext[1])
        half = max(cy - lo[1], hi[1] - cy)
        lo[1], hi[1] = cy - half, cy + half
    if grid == "cube":
        # set_position_range takes ONE resolution for all axes (neighborhood_green.cpp:19): bounds that hug a 600 x 300 x 200 pool
        # give cells of 4.8 x 2.5 x 1.7 r at 128 cells per axis, and a walk of one cutoff around a block of cells visits four
        # times the cells a cubic grid needs.  "cube": every axis gets the longest extent, centred where it was (the bricks of
        # the multi-GPU runs cut the grid in halves, so the fluid stays symmetric about the cuts).  Measured at 10^6 particles:
        # emit 0.71 -> 0.58 ms (profiles/r02_variants.md).
        half = 0.5 * max(h - l for l, h in zip(lo, hi))
        mid = [0.5 * (l + h) for l, h in zip(lo, hi)]
        lo = [m - half for m in mid]; hi = [m + half for m in mid]
    if res_log2 is None and grid == "cube":
        # cubic cells: the resolution whose cell edge is closest to 1.1 kernel widths (4.8 r at 620 r, 3.2 r rather than 6.4 r at 820 r)
        res_log2 = _res_near(max(h - l for l, h in zip(lo, hi)), 4.4 * r)
    if res_log2 is None:
        # per-axis grid bounds (cells need not be cubes); resolution so that the widest cell is about 0.75 x search range
        res_log2 = _res_for(max(h - l for l, h in zip(lo, hi)), 4.5 * r)
    sc = Scene(name=f"dam_break_{nx}x{ny}x{nz}" + ("x2" if blocks == 2 else ""), dims=3, arrays=shuffle_state(arrays, seed),
               min_pos=tuple(lo), max_pos=tuple(hi), res_log2=res_log2,
               basic_pbf=not adaptive, solver_iterations=4, smallest_target_radius=r)
    sc.box_min, sc.box_max = pool_walls(pool_min, pool_max, r, 3)
    return sc


def waterdrop(side=160, r=1.0, jitter=0.05, seed=5, wall_gap=0.0):
    """configs[2]: four radius classes {r, 2^(1/3) r, 2^(2/3) r, 2r} stacked by depth (emulates split/merge output,
    SURVEY 8d-3): finest at the top, coarsest at the bottom, every class seeded on its own lattice of spacing 2 r_c so
    that the rest density is right everywhere -> variable kernel widths, unmirrored pairs, neighbour-count skew.
    `side` = number of finest-class particles that would fit along one axis."""
    L = 2.0 * r * side
    half = 0.5 * L
    parts = []
    for c in range(4):                      # c = 0: top slab (finest) ... c = 3: bottom slab (radius 2r)
        rc = float(np.float32(r) * np.power(np.float32(2.0), np.float32(c) / np.float32(3.0)))
        sp = 2.0 * rc
        nxz = max(int(L / sp), 1)
        ny = max(int((L / 4.0) / sp), 1)
        y_top = half - c * (L / 4.0)
        origin = (-0.5 * sp * (nxz - 1), y_top - rc - sp * (ny - 1), -0.5 * sp * (nxz - 1))
        parts.append(_lattice_state((nxz, ny, nxz), origin, rc, jitter * rc / r if jitter else 0.0, seed + c, 3))
    arrays = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    arrays["index_list"] = np.arange(arrays["position"].shape[0], dtype=np.uint32)
    lo = -(half + 12 * r); hi = half + 12 * r
    sc = Scene(name=f"waterdrop_{side}^3", dims=3, arrays=shuffle_state(arrays, seed), min_pos=(lo,) * 3, max_pos=(hi,) * 3,
               res_log2=_res_for(hi - lo, 6.0 * r), basic_pbf=False, solver_iterations=4, smallest_target_radius=r)
    sc.box_min, sc.box_max = pool_walls((-half - wall_gap,) * 3, (half + wall_gap,) * 3, r, 3)
    return sc
