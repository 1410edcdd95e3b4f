"""apbf_b200 -- B200 (sm_100a) implementation of APBF's per-substep particle hot path.

The product is libapbf_b200.so (C-ABI in include/apbf_b200.h, CUDA sources in apbf_b200/csrc).  This module is the
thin Python front-end used by tests/ and bench.py: it mirrors the reference's operator interface (same class and method
names as source/neighborhood_green.h, incompressibility.h, spread_kernel_width.h, box_collision.h,
velocity_handling.h) on top of device arrays held in torch tensors.  PyTorch is plumbing only (device memory, streams,
torch.distributed); every computation goes through the C-ABI.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import Settings

__all__ = ["Context", "ParticleLists", "neighborhood_green", "neighborhood_binary_search", "incompressibility",
           "spread_kernel_width", "box_collision", "velocity_handling", "algorithms", "Sim", "Settings", "TransferList",
           "update_transfers", "particle_transfer", "NeighborList"]

_HIDDEN = (("position", np.int32, 4), ("velocity", np.float32, 4), ("inverse_mass", np.float32, 1),
           ("radius", np.float32, 1), ("pos_backup", np.int32, 4), ("transferring", np.uint32, 1))
_PER_ID = (("target_radius", np.float32, 1), ("kernel_width", np.float32, 1), ("boundariness", np.float32, 1),
           ("boundary_distance", np.uint32, 1))
FIELDS = (("index_list", np.uint32, 1),) + _HIDDEN + _PER_ID


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("apbf_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _check(ctx, rc):
    if rc != 0:
        msg = _capi.load().apbf_ctx_last_error(ctx.handle) if ctx is not None and ctx.handle else b""
        raise RuntimeError(f"apbf_b200 call failed with status {rc}: {msg.decode() if msg else ''}")


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


_TORCH_DT = {np.dtype(np.int32): "int32", np.dtype(np.float32): "float32", np.dtype(np.uint32): "int32"}


class Context:
    """apbf_ctx: one ordered CUDA stream of work (replaces shader_provider's recording state)."""

    def __init__(self, device=0, stream=None, dims=3, settings=None):
        torch = _torch()
        self.lib = _capi.load()
        self.device = device
        torch.cuda.set_device(device)
        if stream is None:
            stream = torch.cuda.current_stream(device).cuda_stream
        h = C.c_void_p()
        rc = self.lib.apbf_ctx_create(device, C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"apbf_ctx_create failed with status {rc} (no CUDA device? there is no CPU fallback)")
        self.handle = h
        self.dims = dims
        _check(self, self.lib.apbf_ctx_set_dimensions(self.handle, dims))
        self.settings = Settings()
        self.lib.apbf_default_settings(C.byref(self.settings))
        if settings is not None:
            self.set_settings(settings)

    def set_settings(self, s=None, **kw):
        if s is not None:
            self.settings = s
        for k, v in kw.items():
            setattr(self.settings, k, v)
        _check(self, self.lib.apbf_ctx_set_settings(self.handle, C.byref(self.settings)))

    def set_dimensions(self, dims):
        self.dims = dims
        _check(self, self.lib.apbf_ctx_set_dimensions(self.handle, dims))

    def synchronize(self):
        _check(self, self.lib.apbf_ctx_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.apbf_ctx_launch_count(self.handle))

    def set_stream_blocks(self, max_blocks=0):
        """testing aid: cap the hit stream of the pair emit (0 = automatic); an exhausted stream falls back to the two-pass fill"""
        _check(self, self.lib.apbf_ctx_set_stream_blocks(self.handle, int(max_blocks)))

    def set_match_grid_min(self, min_candidates):
        """tuning / testing aid: merge / split matching runs its first rounds grid-wide from this many candidates on"""
        _check(self, self.lib.apbf_ctx_set_match_grid_min(self.handle, int(min_candidates)))

    def set_search_stats(self, enable=True):
        """fused search + spread: also count the pairs of the (never materialised) unpruned list"""
        _check(self, self.lib.apbf_ctx_set_search_stats(self.handle, int(enable)))

    def profile(self, enable=True):
        _check(self, self.lib.apbf_ctx_profile(self.handle, int(enable)))

    def profile_read(self):
        """{category: (milliseconds, spans)} accumulated since profile(True)"""
        out, i = {}, 0
        while True:
            name, ms, calls = C.c_char_p(), C.c_double(), C.c_uint64()
            if self.lib.apbf_ctx_profile_read(self.handle, i, C.byref(name), C.byref(ms), C.byref(calls)) != 0:
                return out
            out[name.value.decode()] = (ms.value, calls.value)
            i += 1

    def device_flags(self):
        f = C.c_uint32()
        _check(self, self.lib.apbf_ctx_device_flags(self.handle, C.byref(f)))
        return f.value

    def list_state(self):
        """which form of the passes the last search / solver iteration selected on the device (apbf_ctx_list_state)"""
        w = (C.c_uint32 * 4)()
        _check(self, self.lib.apbf_ctx_list_state(self.handle, w))
        return {"pairs_unmirrored": int(w[0]), "kernel_widths_uniform": bool(w[1]), "thresholds_uniform": bool(w[2]), "occupied_cells": int(w[3])}

    def close(self):
        if getattr(self, "handle", None):
            self.lib.apbf_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParticleLists:
    """The scene's lists (pbd::particles + pbd::fluid + pbd::neighbors, source/list_definitions.h:9-20) in device memory.
    Every re-orderable array has a second buffer, the copy-on-write target of gpu_list::apply_edit."""

    def __init__(self, ctx, arrays, capacity=None, neighbor_capacity=None):
        torch = _torch()
        self.ctx = ctx
        n = int(np.asarray(arrays["index_list"]).shape[0])
        nh = int(np.asarray(arrays["position"]).reshape(-1, 4).shape[0])
        self.capacity = int(capacity or max(n, nh, 1))
        self.neighbor_capacity = int(neighbor_capacity or 64 * self.capacity)
        dev = torch.device("cuda", ctx.device)
        self.buf = {}
        for name, dt, w in FIELDS:
            src = np.ascontiguousarray(arrays[name], dtype=dt).reshape(-1, w)
            a = torch.zeros((self.capacity, w), dtype=getattr(torch, _TORCH_DT[np.dtype(dt)]), device=dev)
            b = torch.zeros_like(a)
            if src.shape[0]:
                a[:src.shape[0]].copy_(torch.from_numpy(src.view(np.int32) if dt == np.uint32 else src).to(dev))
            self.buf[name] = [a, b]
        self.words = torch.zeros(16, dtype=torch.int32, device=dev)  # [0] index length, [1] hidden length, [2] pair count
        self.words[0] = n
        self.words[1] = nh
        self.pairs = torch.zeros((self.neighbor_capacity, 2), dtype=torch.int32, device=dev)
        self._pair_word = self.words.data_ptr() + 8
        self._forget_pairs()   # torch recycles device addresses: nothing an earlier list left behind applies to this buffer

    def _forget_pairs(self):
        if getattr(self.ctx, "handle", None) and getattr(self, "pairs", None) is not None:
            nb = _capi.Neighbors(self.pairs.data_ptr(), self._pair_word, self.neighbor_capacity)
            self.ctx.lib.apbf_neighbors_release(self.ctx.handle, C.byref(nb))

    def __del__(self):
        try:
            self._forget_pairs()
        except Exception:
            pass

    def use_neighbors(self, other):
        """operators of this object read / write the pair list of `other` (a NeighborList) from now on: set_data(fluid, neighbors)
        of the reference takes the two lists separately"""
        self._forget_pairs()
        self.pairs, self._pair_word, self.neighbor_capacity = other.pairs, other.word.data_ptr(), other.capacity
        self._keep = other
        return self

    def write_pairs(self, pairs):
        """the caller's own pair list: pbd::neighbors::write() followed by the caller's writes.  pairs: host array [P, 2]"""
        torch = _torch()
        a = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        assert a.shape[0] <= self.neighbor_capacity
        if a.shape[0]:
            self.pairs[: a.shape[0]].copy_(torch.from_numpy(a.view(np.int32)).to(self.pairs.device))
        word = (self.words[2:3] if self._pair_word == self.words.data_ptr() + 8 else self._keep.word)
        word.fill_(int(a.shape[0]))
        nb = self.neighbors()
        _check(self.ctx, self.ctx.lib.apbf_neighbors_invalidate(self.ctx.handle, C.byref(nb)))

    # ---- C views --------------------------------------------------------------------------------------------------
    def _arr(self, name):
        a, b = self.buf[name]
        return _capi.Array(a.data_ptr(), b.data_ptr())

    def fluid(self):
        p = _capi.Particles()
        p.index_list = self._arr("index_list")
        p.length = self.words.data_ptr()
        p.capacity = self.capacity
        p.hidden_length = self.words.data_ptr() + 4
        p.hidden_capacity = self.capacity
        for name, _, _ in _HIDDEN:
            setattr(p, name, self._arr(name))
        f = _capi.Fluid()
        f.particle = p
        for name, _, _ in _PER_ID:
            setattr(f, name, self._arr(name))
        return f

    def neighbors(self):
        return _capi.Neighbors(self.pairs.data_ptr(), self._pair_word, self.neighbor_capacity)

    def swap(self):
        """after a search: the reorder_out buffers hold the lists"""
        for v in self.buf.values():
            v.reverse()

    # ---- read-back (gpu_list::read, debugging only) --------------------------------------------------------------
    def length(self):
        return int(self.words[0].item())

    def pair_count(self):
        if self._pair_word != self.words.data_ptr() + 8:
            return int(self._keep.word.item())
        return int(self.words[2].item())

    def read(self, name):
        _, dt, w = next(f for f in FIELDS if f[0] == name)
        n = self.length() if name not in [h[0] for h in _HIDDEN] else int(self.words[1].item())
        a = self.buf[name][0][:n].cpu().numpy()
        a = a.view(np.uint32) if dt == np.uint32 else a
        return a.reshape(-1, w) if w > 1 else a.reshape(-1)

    def read_all(self):
        return {name: self.read(name) for name, _, _ in FIELDS}

    def write(self, name, values):
        """overwrite the first len(values) entries of one list from a host array (tests)"""
        torch = _torch()
        _, dt, w = next(f for f in FIELDS if f[0] == name)
        a = np.ascontiguousarray(values, dtype=dt).reshape(-1, w)
        t = torch.from_numpy(a.view(np.int32) if dt == np.uint32 else a).to(self.words.device)
        self.buf[name][0][: a.shape[0]].view(-1, w).copy_(t)

    def read_pairs(self):
        return self.pairs[:self.pair_count()].cpu().numpy().view(np.uint32)


class NeighborList:
    """a pbd::neighbors list of its own (gpu_list<8> + length word), e.g. a second list over the same particles"""

    def __init__(self, ctx, capacity, pairs=None):
        torch = _torch()
        self.ctx, self.capacity = ctx, int(capacity)
        dev = torch.device("cuda", ctx.device)
        self.pairs = torch.zeros((self.capacity, 2), dtype=torch.int32, device=dev)
        self.word = torch.zeros(4, dtype=torch.int32, device=dev)
        nb = self.c()
        ctx.lib.apbf_neighbors_release(ctx.handle, C.byref(nb))
        if pairs is not None:
            a = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
            self.pairs[: a.shape[0]].copy_(torch.from_numpy(a.view(np.int32)).to(dev))
            self.word[0] = a.shape[0]

    def c(self):
        return _capi.Neighbors(self.pairs.data_ptr(), self.word.data_ptr(), self.capacity)

    def read(self):
        return self.pairs[: int(self.word[0].item())].cpu().numpy().view(np.uint32)

    def __del__(self):
        try:
            if getattr(self.ctx, "handle", None):
                nb = self.c()
                self.ctx.lib.apbf_neighbors_release(self.ctx.handle, C.byref(nb))
        except Exception:
            pass


class TransferList:
    """pbd::transfers (source/list_definitions.h:16-18): rows (source, target, time_left) in device memory; source and target
    index the hidden particle list.  Two buffers per list, like ParticleLists."""

    def __init__(self, ctx, capacity, source=(), target=(), time_left=()):
        torch = _torch()
        self.ctx, self.capacity = ctx, int(capacity)
        dev = torch.device("cuda", ctx.device)
        self.buf = {k: [torch.zeros(self.capacity, dtype=dt, device=dev), torch.zeros(self.capacity, dtype=dt, device=dev)]
                    for k, dt in (("source", torch.int32), ("target", torch.int32), ("time_left", torch.float32))}
        n = len(source)
        if n:
            self.buf["source"][0][:n] = torch.from_numpy(np.asarray(source, np.uint32).view(np.int32)).to(dev)
            self.buf["target"][0][:n] = torch.from_numpy(np.asarray(target, np.uint32).view(np.int32)).to(dev)
            self.buf["time_left"][0][:n] = torch.from_numpy(np.asarray(time_left, np.float32)).to(dev)
        self.word = torch.tensor([n], dtype=torch.int32, device=dev)

    def c(self):
        t = _capi.Transfers()
        for k in ("source", "target", "time_left"):
            a, b = self.buf[k]
            setattr(t, k, _capi.Array(a.data_ptr(), b.data_ptr()))
        t.length, t.capacity = self.word.data_ptr(), self.capacity
        return t

    def swap(self):
        for v in self.buf.values():
            v.reverse()

    def length(self):
        return int(self.word.item())

    def rows(self):
        n = self.length()
        return (self.buf["source"][0][:n].cpu().numpy().view(np.uint32), self.buf["target"][0][:n].cpu().numpy().view(np.uint32),
                self.buf["time_left"][0][:n].cpu().numpy())


class _Operator:
    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib


class neighborhood_green(_Operator):
    """pbd::neighborhood_green (source/neighborhood_green.h:8-24)"""

    def set_data(self, lists):
        self.lists = lists
        return self

    def set_range_scale(self, scale):
        self.scale = float(scale)
        return self

    def set_position_range(self, min_pos, max_pos, resolution_log2):
        self.min_pos, self.max_pos, self.res = tuple(min_pos), tuple(max_pos), int(resolution_log2)
        return self

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        rng = fl.kernel_width  # pool.cpp:45: the range list is fluid.kernel_width
        dbg, keep = None, {}
        if debug:
            dev = L.words.device
            n_cells = 1 << (self.res * self.ctx.dims)
            keep = dict(sorted_key=torch.zeros(L.capacity, dtype=torch.int32, device=dev),
                        sorted_index=torch.zeros(L.capacity, dtype=torch.int32, device=dev),
                        cell_start=torch.zeros(n_cells, dtype=torch.int32, device=dev),
                        cell_end=torch.zeros(n_cells, dtype=torch.int32, device=dev),
                        pair_offsets=torch.zeros(L.capacity + 1, dtype=torch.int32, device=dev))
            dbg = _capi.SearchDebug(keep["sorted_key"].data_ptr(), keep["sorted_index"].data_ptr(),
                                    keep["cell_start"].data_ptr(), keep["cell_end"].data_ptr())
            dbg.pair_offsets = keep["pair_offsets"].data_ptr()
        _check(self.ctx, self.lib.apbf_neighborhood_green_apply(
            self.ctx.handle, C.byref(fl), C.byref(rng), C.byref(nb), self.scale, _f3(self.min_pos), _f3(self.max_pos),
            self.res, C.byref(dbg) if dbg else None))
        L.swap()
        if debug:
            nh = int(L.words[1].item())
            out = {k: v.cpu().numpy().view(np.uint32) for k, v in keep.items()}
            out["sorted_key"] = out["sorted_key"][:nh]
            out["sorted_index"] = out["sorted_index"][:nh]
            out["pair_offsets"] = out["pair_offsets"][:L.length() + 1]
            return out


class neighborhood_green_spread(neighborhood_green):
    """neighborhood_green::apply() + spread_kernel_width::apply() in one pass (source/pool.cpp:83-89); not a class of the
    reference: what the whole-scene substep runs for adaptive kernel widths (apbf_neighborhood_green_spread_apply)."""

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        kwfx = torch.zeros(L.capacity, dtype=torch.int32, device=L.words.device) if debug else None
        _check(self.ctx, self.lib.apbf_neighborhood_green_spread_apply(
            self.ctx.handle, C.byref(fl), C.byref(nb), self.scale, _f3(self.min_pos), _f3(self.max_pos), self.res, None,
            kwfx.data_ptr() if debug else None))
        L.swap()
        if debug:
            return kwfx[:L.length()].cpu().numpy().view(np.uint32)


class neighborhood_binary_search(_Operator):
    """pbd::neighborhood_binary_search (source/neighborhood_binary_search.h:8-21)"""

    def set_data(self, lists):
        self.lists = lists
        return self

    def set_range_scale(self, scale):
        self.scale = float(scale)
        return self

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        rng = fl.kernel_width
        dbg, keep = None, {}
        if debug:
            dev = L.words.device
            keep = {k: torch.zeros(L.capacity, dtype=torch.int32, device=dev) for k in ("sorted_index", "code0", "code1", "code2")}
            dbg = _capi.SearchDebug()
            dbg.sorted_index = keep["sorted_index"].data_ptr()
            dbg.code[0], dbg.code[1], dbg.code[2] = keep["code0"].data_ptr(), keep["code1"].data_ptr(), keep["code2"].data_ptr()
        _check(self.ctx, self.lib.apbf_neighborhood_binary_search_apply(
            self.ctx.handle, C.byref(fl), C.byref(rng), C.byref(nb), self.scale, C.byref(dbg) if dbg else None))
        L.swap()
        if debug:
            n = L.length()
            return {k: v.cpu().numpy().view(np.uint32)[:n] for k, v in keep.items()}


class neighborhood_binary_search_spread(neighborhood_binary_search):
    """neighborhood_binary_search::apply() + spread_kernel_width::apply() in one pass (pool.cpp:83-89 with NEIGHBORHOOD_TYPE 3);
    not a class of the reference: what the whole-scene substep runs for adaptive widths with the binary search"""

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        kwfx = torch.zeros(L.capacity, dtype=torch.int32, device=L.words.device) if debug else None
        _check(self.ctx, self.lib.apbf_neighborhood_binary_search_spread_apply(
            self.ctx.handle, C.byref(fl), C.byref(nb), self.scale, None, kwfx.data_ptr() if debug else None))
        L.swap()
        if debug:
            return kwfx[:L.length()].cpu().numpy().view(np.uint32)


class incompressibility(_Operator):
    """pbd::incompressibility (source/incompressibility.h:8-18)"""

    def set_data(self, lists):
        self.lists = lists
        return self

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        lam = inc = None
        if debug:
            lam = torch.zeros(L.capacity, dtype=torch.float32, device=L.words.device)
            inc = torch.zeros((L.capacity, 8), dtype=torch.int32, device=L.words.device)
        _check(self.ctx, self.lib.apbf_incompressibility_apply(
            self.ctx.handle, C.byref(fl), C.byref(nb), lam.data_ptr() if debug else None, inc.data_ptr() if debug else None))
        if debug:
            n = L.length()
            inc = inc[:n].cpu().numpy()
            return dict(lam=lam[:n].cpu().numpy(), grad_sum=inc[:, 0:3].copy(), density=inc[:, 3].view(np.uint32).copy(),
                        sq_grad_sum=inc[:, 4].view(np.uint32).copy())


class spread_kernel_width(_Operator):
    """pbd::spread_kernel_width (source/spread_kernel_width.h:7-17)"""

    def set_data(self, lists):
        self.lists = lists
        return self

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        kwfx = torch.zeros(L.capacity, dtype=torch.int32, device=L.words.device) if debug else None
        _check(self.ctx, self.lib.apbf_spread_kernel_width_apply(self.ctx.handle, C.byref(fl), C.byref(nb),
                                                                 kwfx.data_ptr() if debug else None))
        if debug:
            return kwfx[:L.length()].cpu().numpy().view(np.uint32)


class update_transfers(_Operator):
    """pbd::update_transfers (source/update_transfers.h, update_transfers.cpp:14-70): boundary-distance flood-fill step, nearest
    neighbour, target radius, boundary-distance decay, boundariness threshold (find_split_and_merge_1/2/3.comp); with a transfer
    list also the merge / split decisions and the start of the splits, as mMerge / mSplit of the context's settings say"""

    def set_data(self, lists, transfers=None):
        self.lists, self.transfers = lists, transfers
        return self

    def set_split_duration(self, split_duration):
        """settings::splitDuration (a host-side setting of the reference, update_transfers.cpp:63)"""
        self.split_duration = float(split_duration)
        return self

    def apply(self, debug=False):
        torch = _torch()
        L = self.lists
        fl, nb = L.fluid(), L.neighbors()
        nearest = torch.zeros(L.capacity, dtype=torch.int32, device=L.words.device) if debug else None
        n = L.length()
        if self.transfers is None:
            _check(self.ctx, self.lib.apbf_update_transfers_apply(self.ctx.handle, C.byref(fl), C.byref(nb),
                                                                  nearest.data_ptr() if debug else None))
        else:
            tr = self.transfers.c()
            _check(self.ctx, self.lib.apbf_update_transfers_split_merge_apply(
                self.ctx.handle, C.byref(fl), C.byref(nb), C.byref(tr), getattr(self, "split_duration", 0.0),
                nearest.data_ptr() if debug else None))
        if debug:
            return nearest[:n].cpu().numpy().view(np.uint32)


class particle_transfer(_Operator):
    """pbd::particle_transfer (source/particle_transfer.h, particle_transfer.cpp:10-28)"""

    def set_data(self, lists, transfers):
        self.lists, self.transfers = lists, transfers
        return self

    def apply(self, delta_time):
        fl, tr = self.lists.fluid(), self.transfers.c()
        _check(self.ctx, self.lib.apbf_particle_transfer_apply(self.ctx.handle, C.byref(fl), C.byref(tr), float(delta_time), None))
        self.lists.swap()
        self.transfers.swap()


def kernel_width_from_boundary_distance(ctx, lists):
    """pool.cpp:77-80: shader_provider::uint_to_float_with_indexed_lower_bound on boundary_distance -> kernel_width"""
    fl = lists.fluid()
    _check(ctx, ctx.lib.apbf_kernel_width_from_boundary_distance(ctx.handle, C.byref(fl)))


def _fmt_g(x):
    """a float the way `std::ostream << float` prints it by default: printf("%g"), six significant digits"""
    return "%g" % float(np.float32(x))


def save_particle_info(lists, folder="particle_data"):
    """pbd::save_particle_info::set_data(fluid, neighbors, transfers).apply() on device lists (reads them back first)"""
    return write_particle_info(lists.read_all(), lists.read_pairs(), folder)


def write_particle_info(a, pairs, folder="particle_data"):
    """pbd::save_particle_info::apply() (source/save_particle_info.cpp:21-130): the reference's on-disk particle dump -- six
    ';'-separated text files and data.csv, sorted by the distance to (0, 10, -60).  Host-side only: lets a run of this library
    be diffed against a run of the reference.  `a`: the lists as host arrays (FIELDS), `pairs`: [P, 2].
    Returns the folder."""
    import os
    f32 = np.float32
    idx = np.asarray(a["index_list"]).astype(np.int64)
    n = len(idx)
    pos = (np.asarray(a["position"])[idx, :3].astype(f32) / f32(262144.0)).astype(f32)
    pairs = np.asarray(pairs).reshape(-1, 2)
    nbr_count = np.bincount(pairs[:, 0].astype(np.int64), minlength=n)[:n] if len(pairs) else np.zeros(n, np.int64)
    bdr = (np.asarray(a["boundary_distance"]).astype(f32) / f32(262144.0)).astype(f32)
    d = pos - np.array([0.0, 10.0, -60.0], f32)                      # centerPos, :49
    center = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + d[:, 2] * d[:, 2]).astype(f32)
    radius, inv_mass = np.asarray(a["radius"])[idx], np.asarray(a["inverse_mass"])[idx]
    os.makedirs(folder, exist_ok=True)
    for name, v, ints in (("centerDist.txt", center, False), ("radius.txt", radius, False), ("neighborCount.txt", nbr_count, True),
                          ("kernelWidth.txt", a["kernel_width"], False), ("targetRadius.txt", a["target_radius"], False),
                          ("boundaryDistance.txt", bdr, False)):
        with open(os.path.join(folder, name), "w") as f:
            f.write("".join((str(int(x)) if ints else _fmt_g(x)) + ";" for x in v))
    order = np.argsort(center, kind="stable")                        # std::sort: the order of equal distances is unspecified
    with open(os.path.join(folder, "data.csv"), "w") as f:
        f.write("center distance,boundary distance,kernel width,neighbor count,radius,target radius,inverse mass,x,y,z\n")
        for k in order:
            f.write(",".join([_fmt_g(center[k]), _fmt_g(bdr[k]), _fmt_g(a["kernel_width"][k]), str(int(nbr_count[k])), _fmt_g(radius[k]),
                              _fmt_g(a["target_radius"][k]), _fmt_g(inv_mass[k]), _fmt_g(pos[k, 0]), _fmt_g(pos[k, 1]), _fmt_g(pos[k, 2])]) + "\n")
    return folder


class box_collision(_Operator):
    """pbd::box_collision (source/box_collision.h:8-18)"""

    def set_data(self, lists, box_min, box_max):
        torch = _torch()
        self.lists = lists
        dev = lists.words.device
        self.box_min = torch.from_numpy(np.ascontiguousarray(box_min, np.float32).reshape(-1, 4)).to(dev)
        self.box_max = torch.from_numpy(np.ascontiguousarray(box_max, np.float32).reshape(-1, 4)).to(dev)
        return self

    def apply(self):
        fl = self.lists.fluid()
        _check(self.ctx, self.lib.apbf_box_collision_apply(self.ctx.handle, C.byref(fl.particle), self.box_min.data_ptr(),
                                                           self.box_max.data_ptr(), self.box_min.shape[0]))


class velocity_handling(_Operator):
    """pbd::velocity_handling (source/velocity_handling.h:7-19)"""

    def set_data(self, lists):
        self.lists = lists
        self.last_dt = 1.0
        self.accel = (0.0, 0.0, 0.0)
        return self

    def set_acceleration(self, accel=(0.0, 0.0, 0.0)):
        self.accel = tuple(accel)
        return self

    def apply(self, dt):
        fl = self.lists.fluid()
        _check(self.ctx, self.lib.apbf_velocity_handling_apply(self.ctx.handle, C.byref(fl.particle), float(dt),
                                                               float(self.last_dt), _f3(self.accel)))
        if dt != 0.0:
            self.last_dt = dt


class algorithms:
    """pbd::algorithms (source/algorithms.h:12-17) on torch int32 tensors holding u32 values."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib

    def sort(self, keys, values, count, max_count, out_keys, out_values, upper_bound=0xFFFFFFFF):
        _check(self.ctx, self.lib.apbf_sort(self.ctx.handle, keys.data_ptr(), values.data_ptr(), count.data_ptr(), max_count,
                                            out_keys.data_ptr(), out_values.data_ptr(), upper_bound))

    def prefix_sum(self, values, count, max_count, result=None):
        result = values if result is None else result
        _check(self.ctx, self.lib.apbf_prefix_sum(self.ctx.handle, values.data_ptr(), count.data_ptr(), max_count, result.data_ptr()))

    def sort_calculate_needed_helper_list_length(self, n):
        return self.lib.apbf_sort_calculate_needed_helper_list_length(n)

    def prefix_sum_calculate_needed_helper_list_length(self, n):
        return self.lib.apbf_prefix_sum_calculate_needed_helper_list_length(n)


class Sim:
    """apbf_sim: the lists of one scene resident in HBM and pool::update (source/pool.cpp:67-106) over them.
    upload()/download() move the lists between host memory and the device (the end-to-end path)."""

    def __init__(self, ctx, scene, capacity=None, neighbor_capacity=None, use_binary_search=False, integrate=False,
                 dt=1.0 / 60.0, accel=(0.0, -10.0, 0.0), solver_iterations=None, basic_pbf=None, update_transfers=False,
                 transfers=False, transfer_capacity=0, split_duration=0.0):
        self.ctx, self.lib = ctx, ctx.lib
        self.capacity = int(capacity or scene.n)
        self.neighbor_capacity = int(neighbor_capacity or 40 * self.capacity)
        cfg = _capi.SimConfig()
        cfg.particle_capacity, cfg.neighbor_capacity = self.capacity, self.neighbor_capacity
        cfg.dims = scene.dims
        cfg.basic_pbf = int(scene.basic_pbf if basic_pbf is None else basic_pbf)
        cfg.solver_iterations = int(scene.solver_iterations if solver_iterations is None else solver_iterations)
        cfg.use_binary_search, cfg.integrate, cfg.dt = int(use_binary_search), int(integrate), dt
        cfg.accel, cfg.min_pos, cfg.max_pos = _f3(accel), _f3(scene.min_pos), _f3(scene.max_pos)
        cfg.res_log2 = scene.res_log2
        cfg.update_transfers = int(update_transfers)
        cfg.transfers, cfg.transfer_capacity, cfg.split_duration = int(transfers), int(transfer_capacity), float(split_duration)
        self.transfer_capacity = int(transfer_capacity or self.capacity) if transfers else 0
        bmin = np.ascontiguousarray(scene.box_min, np.float32).reshape(-1, 4)
        bmax = np.ascontiguousarray(scene.box_max, np.float32).reshape(-1, 4)
        cfg.n_boxes = bmin.shape[0]
        cfg.box_min4_host, cfg.box_max4_host = bmin.ctypes.data, bmax.ctypes.data
        h = C.c_void_p()
        _check(ctx, self.lib.apbf_sim_create(ctx.handle, C.byref(cfg), C.byref(h)))
        self.handle = h

    @staticmethod
    def host_state(arrays, n=None):
        """arrays: name -> numpy array or (pinned) torch CPU tensor"""
        hs = _capi.HostState()
        hs.n = int(n if n is not None else len(arrays["position"]))
        for name, _, _ in FIELDS:
            a = arrays.get(name)
            if a is None:
                continue
            ptr = a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
            setattr(hs, name, ptr)
        return hs

    def upload(self, arrays, n=None):
        hs = self.host_state(arrays, n)
        _check(self.ctx, self.lib.apbf_sim_upload(self.handle, C.byref(hs)))

    def download(self, arrays):
        hs = self.host_state(arrays, 0)
        _check(self.ctx, self.lib.apbf_sim_download(self.handle, C.byref(hs)))
        return hs.n

    def substep(self, n=1):
        _check(self.ctx, self.lib.apbf_sim_substep(self.handle, n))

    def step_host(self, arrays_in, n, arrays_out=None):
        """host buffers in -> one substep -> host buffers out with the copies overlapped with the work (apbf_sim_step_host);
        returns the number of particles that came back"""
        hi = self.host_state(arrays_in, n)
        ho = self.host_state(arrays_out if arrays_out is not None else arrays_in, 0)
        _check(self.ctx, self.lib.apbf_sim_step_host(self.handle, C.byref(hi), C.byref(ho)))
        return ho.n

    def set_graphs(self, enable):
        """replay substeps as captured CUDA graphs (on by default, see apbf_sim_set_graphs)"""
        _check(self.ctx, self.lib.apbf_sim_set_graphs(self.handle, 1 if enable else 0))

    def graph_replays(self):
        c = C.c_uint64()
        _check(self.ctx, self.lib.apbf_sim_graph_replays(self.handle, C.byref(c)))
        return c.value

    def stats(self):
        w = (C.c_uint32 * 4)()
        _check(self.ctx, self.lib.apbf_sim_stats(self.handle, w))
        return dict(n=w[0], pairs_searched=w[1], pairs_kept=w[2], pairs_unmirrored=w[3])

    def download_transfers(self):
        """(source, target, time_left) rows of the scene's transfer list (synchronises)"""
        src, tgt = np.zeros(self.transfer_capacity, np.uint32), np.zeros(self.transfer_capacity, np.uint32)
        ttl, n = np.zeros(self.transfer_capacity, np.float32), C.c_uint32()
        _check(self.ctx, self.lib.apbf_sim_download_transfers(self.handle, C.byref(n), src.ctypes.data, tgt.ctypes.data, ttl.ctypes.data))
        return src[:n.value], tgt[:n.value], ttl[:n.value]

    def neighbor_count(self):
        c = C.c_uint32()
        _check(self.ctx, self.lib.apbf_sim_neighbor_count(self.handle, C.byref(c)))
        return c.value

    def read_pairs(self):
        """the scene's pair list as (id, idN) rows on the host: apbf_sim_neighbors() writes the public list on demand"""
        nb = _capi.Neighbors()
        _check(self.ctx, self.lib.apbf_sim_neighbors(self.handle, C.byref(nb)))
        n = min(self.neighbor_count(), self.neighbor_capacity)
        out = np.zeros((n, 2), np.uint32)
        if n:
            _check(self.ctx, self.lib.apbf_copy_bytes_to_host(self.ctx.handle, nb.pairs, out.ctypes.data, 8 * n))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.apbf_sim_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def empty_host_arrays(n):
    return {name: np.zeros((n, w) if w > 1 else (n,), dt) for name, dt, w in FIELDS}
