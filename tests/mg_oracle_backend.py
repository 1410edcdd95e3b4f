"""A rank backend for apbf_b200.multi_gpu.SlabDomain made of numpy and the CPU oracle (TEST INFRASTRUCTURE).

It lets the world_size-2 gloo test run the real host-side protocol (routing, send lists, ghost slots, the order of the
exchanges) without a GPU: every device step of CudaRankBackend has a numpy/oracle twin here.  The oracle keeps the
reference's scatter formulation (pairs push onto their second particle), so a ghost's pair list is complete here, whereas
the CUDA path only keeps a ghost's unmirrored pairs -- the exchanged quantities and their order are the same.
"""
import numpy as np
import torch

from apbf_b200.multi_gpu import HALO, KW, LAMBDA, P4, POS
from oracle import oracle as orc

FIELDS = [f[0] for f in orc.State.FIELDS if f[0] != "index_list"]


def _deinterleave(key, res, dims):
    cells = np.zeros((len(key), 3), np.uint32)
    for i in range(res):
        for d in range(dims):
            cells[:, d] |= ((key >> np.uint32(i * dims + d)) & np.uint32(1)) << np.uint32(i)
    return cells


class OracleRankBackend:
    def __init__(self, arrays, scene, settings, rank, world, halo_range, adaptive, cap_pairs):
        self.a = {k: np.ascontiguousarray(arrays[k]).copy() for k in FIELDS}
        self.sc, self.s, self.rank, self.world = scene, settings, rank, world
        self.adaptive, self.cap_pairs = adaptive, cap_pairs
        self.n_own = self.n_tot = len(self.a["position"])
        self.gid_base = 0
        res, dims = scene.res_log2, scene.dims
        self.lw = int(np.log2(world))
        cells = 1 << res
        self.halo = [int(np.ceil(halo_range / ((scene.max_pos[d] - scene.min_pos[d]) / cells))) + 1 if d < dims else 0 for d in range(3)]
        self.bricks = []
        for r in range(world):
            lo, size = [0, 0, 0], [cells, cells, cells if dims == 3 else 1]
            for b in range(self.lw):
                axis = dims - 1 - (b % dims)
                size[axis] >>= 1
                if (r >> (self.lw - 1 - b)) & 1:
                    lo[axis] += size[axis]
            self.bricks.append((lo, [lo[d] + size[d] - 1 for d in range(3)]))

    # ---- helpers ------------------------------------------------------------------------------------------------------------
    def _state(self, n):
        arrays = {k: self.a[k][:n] for k in FIELDS}
        arrays["index_list"] = np.arange(n, dtype=np.uint32)
        return orc.State(**arrays)

    def _store(self, st, n):
        for k in FIELDS:
            self.a[k] = np.concatenate([getattr(st, k)[:n], self.a[k][n:]]) if len(self.a[k]) > n else getattr(st, k)[:n].copy()

    def _keys(self, n):
        sc = self.sc
        return orc.position_hash(self.a["position"][:n], sc.min_pos, sc.max_pos, sc.res_log2, sc.dims)

    def n_owned(self):
        return self.n_own

    def set_counts(self, n_owned, n_total, gid_base):
        self.n_own, self.n_tot, self.gid_base = n_owned, n_total, gid_base
        for k in FIELDS:
            a = self.a[k]
            if len(a) > n_total:
                self.a[k] = a[:n_total].copy()
            elif len(a) < n_total:
                self.a[k] = np.concatenate([a, np.zeros((n_total - len(a),) + a.shape[1:], a.dtype)])

    # ---- device steps -----------------------------------------------------------------------------------------------------------
    def integrate(self):
        st = self._state(self.n_own)
        orc.velocity_handling(st, 1.0 / 60.0, (0.0, -10.0, 0.0))
        self._store(st, self.n_own)

    def route(self):
        key = self._keys(self.n_own)
        bits = self.sc.res_log2 * self.sc.dims
        dest = (key >> np.uint32(bits - self.lw)).astype(np.int64) if self.lw else np.zeros(self.n_own, np.int64)
        order = np.argsort(dest, kind="stable")
        for k in FIELDS:
            self.a[k] = self.a[k][: self.n_own][order].copy()
        return np.bincount(dest, minlength=self.world).tolist()

    def pack_state(self, first, count):
        cols = [self.a[k][first:first + count].reshape(count, -1).view(np.int32) for k in FIELDS]
        return torch.from_numpy(np.ascontiguousarray(np.concatenate(cols + [np.zeros((count, 1), np.int32)], axis=1)))

    def assemble(self, stay_first, stay_count, arrivals, me):
        parts, placed = [], False
        stay = {k: self.a[k][stay_first:stay_first + stay_count] for k in FIELDS}
        for r, buf in arrivals + [(None, None)]:
            if not placed and (r is None or r > me):
                parts.append(stay)
                placed = True
            if r is not None:
                w = buf.numpy()
                rec, col = {}, 0
                for k in FIELDS:
                    width = self.a[k].reshape(len(self.a[k]), -1).shape[1] if len(self.a[k]) else (4 if self.a[k].ndim == 2 else 1)
                    rec[k] = w[:, col:col + width].copy().view(self.a[k].dtype).reshape((len(w),) + self.a[k].shape[1:])
                    col += width
                parts.append(rec)
        for k in FIELDS:
            self.a[k] = np.concatenate([p[k] for p in parts])
        self.n_own = self.n_tot = len(self.a["position"])

    def halo_lists(self):
        cells = _deinterleave(self._keys(self.n_own), self.sc.res_log2, self.sc.dims).astype(np.int64)
        grid = 1 << self.sc.res_log2
        self.send_ids = {}
        counts = [0] * self.world
        for r in range(self.world):
            if r == self.rank:
                continue
            lo, hi = self.bricks[r]
            inside = np.ones(self.n_own, bool)
            for d in range(self.sc.dims):
                inside &= (cells[:, d] >= max(lo[d] - self.halo[d], 0)) & (cells[:, d] <= min(hi[d] + self.halo[d], grid - 1))
            self.send_ids[r] = np.nonzero(inside)[0]
            counts[r] = int(inside.sum())
        return counts

    def begin_ghosts(self, n_ghost):
        self.ghost_first = self.n_own
        self.ghost_ids = np.arange(self.n_own, self.n_own + n_ghost)

    def pack(self, what, dest):
        ids = self.send_ids[dest]
        if what == HALO:
            cols = [self.a["position"][ids].view(np.int32)] + [self.a[k][ids].reshape(-1, 1).view(np.int32) for k in ("inverse_mass", "radius", "kernel_width", "target_radius")]
            cols += [self.a["boundary_distance"][ids].reshape(-1, 1).view(np.int32), np.zeros((len(ids), 3), np.int32)]
        elif what == KW:
            cols = [self.a["kernel_width"][ids].reshape(-1, 1).view(np.int32)]
        elif what in (P4, POS):
            cols = [self.a["position"][ids].view(np.int32)]
        else:
            cols = [self.lam[ids].reshape(-1, 1).view(np.int32)]
        return torch.from_numpy(np.ascontiguousarray(np.concatenate(cols, axis=1)))

    def unpack(self, what, ghost_offset, count, buf):
        ids = self.ghost_ids[ghost_offset:ghost_offset + count]
        w = buf.numpy()
        if what == HALO:
            self.a["position"][ids] = w[:, :4]
            for i, k in enumerate(("inverse_mass", "radius", "kernel_width", "target_radius")):
                self.a[k][ids] = w[:, 4 + i].copy().view(np.float32)
            self.a["boundary_distance"][ids] = w[:, 8].copy().view(np.uint32)
        elif what == KW:
            self.a["kernel_width"][ids] = w[:, 0].copy().view(np.float32)
        elif what in (P4, POS):
            self.a["position"][ids] = w
        else:
            self.lam[ids] = w[:, 0].copy().view(np.float32)

    def search(self):
        sc = self.sc
        st = self._state(self.n_tot)
        scale = 1.5 if self.adaptive else 1.0
        self.pairs, aux = orc.green_apply(st, self.s, sc.dims, scale, sc.min_pos, sc.max_pos, sc.res_log2, self.cap_pairs, want_aux=True)
        self._store(st, self.n_tot)
        self.inv = np.zeros(self.n_tot, np.int64)
        self.inv[aux["sorted_index"].astype(np.int64)] = np.arange(self.n_tot)
        # the oracle sorts owned and ghost particles into one sequence: remember who is owned
        self.owned_mask = np.zeros(self.n_tot, bool)
        self.owned_mask[self.inv[: self.n_own]] = True

    def remap_after_search(self, send_counts, n_ghost):
        for r in self.send_ids:
            self.send_ids[r] = self.inv[self.send_ids[r]]
        self.ghost_ids = self.inv[self.ghost_ids]

    def spread(self):
        st = self._state(self.n_tot)
        self.pairs, _ = orc.spread_kernel_width_apply(st, self.s, self.pairs)
        self._store(st, self.n_tot)

    def prepare(self):
        pass

    def iter_begin(self, it):
        st = self._state(self.n_tot)
        orc.box_collision(st, self.sc.box_min, self.sc.box_max)   # scenes of this test keep the fluid away from the walls
        self._store(st, self.n_tot)

    def density_lambda(self):
        self.st = self._state(self.n_tot)
        self.incomp, self.grad4, self.lam = orc.incompressibility_passes_012(self.st, self.s, self.sc.dims, self.pairs)

    def apply_delta(self):
        orc.incompressibility_pass_3(self.st, self.s, self.sc.dims, self.pairs, self.incomp, self.grad4, self.lam)
        self._store(self.st, self.n_tot)

    def final_commit(self):
        # owned particles first again (the CUDA path keeps them dense by construction)
        order = np.concatenate([np.nonzero(self.owned_mask)[0], np.nonzero(~self.owned_mask)[0]])
        for k in FIELDS:
            self.a[k] = self.a[k][order].copy()
        inv = np.zeros(self.n_tot, np.int64)
        inv[order] = np.arange(self.n_tot)
        self.pairs = inv[self.pairs.astype(np.int64)].astype(np.uint32)        # the pair list follows (update_transfers reads it)
        for r in self.send_ids:
            self.send_ids[r] = inv[self.send_ids[r]]
        self.ghost_ids = inv[self.ghost_ids]

    def kernel_width_from_boundary_distance(self):
        st = self._state(self.n_own)
        orc.kernel_width_from_boundary_distance(st, self.s)
        self._store(st, self.n_own)

    def update_transfers(self):
        st = self._state(self.n_tot)
        order = np.argsort(self.pairs[:, 0], kind="stable")                    # grouped by id again, discovery order inside
        orc.update_transfers_apply(st, self.s, self.pairs[order])
        self._store(st, self.n_tot)

    def owned_arrays(self):
        return {k: self.a[k][: self.n_own] for k in FIELDS}
