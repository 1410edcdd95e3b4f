"""ctypes front-end of the CPU oracle (oracle/apbf_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from apbf_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libapbf_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("apbf_oracle.c", "apbf_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libapbf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Settings(C.Structure):
    """apbf_settings, shaders/cpu_gpu_shared_config.h:74-94."""
    _fields_ = [(n, C.c_int) for n in (
        "mHeightKernelId", "mGradientKernelId", "mMerge", "mSplit", "mBaseKernelWidthOnTargetRadius",
        "mBaseKernelWidthOnBoundaryDistance", "mUpdateTargetRadius", "mUpdateBoundariness",
        "mNeighborListSorted", "mBoundarinessCalculationMethod")] + [(n, C.c_float) for n in (
        "mBoundarinessAdaptionSpeed", "mKernelWidthAdaptionSpeed", "mBoundarinessSelfGradLengthFactor",
        "mBoundarinessUnderpressureFactor", "mMergeDuration", "mSmallestTargetRadius", "mTargetRadiusOffset",
        "mTargetRadiusScaleFactor")]


class _State(C.Structure):
    _fields_ = [("n_hidden", C.c_uint32), ("n", C.c_uint32)] + [(n, C.c_void_p) for n in (
        "index_list", "position", "velocity", "inverse_mass", "radius", "pos_backup", "transferring",
        "target_radius", "kernel_width", "boundariness", "boundary_distance")]


class _SubstepParams(C.Structure):
    _fields_ = [("dims", C.c_int), ("basic_pbf", C.c_int), ("solver_iterations", C.c_int),
                ("use_binary_search", C.c_int), ("integrate", C.c_int), ("dt", C.c_float),
                ("accel", C.c_float * 3), ("min_pos", C.c_float * 3), ("max_pos", C.c_float * 3),
                ("res_log2", C.c_uint32), ("n_boxes", C.c_uint32), ("box_min4", C.c_void_p), ("box_max4", C.c_void_p),
                ("update_transfers", C.c_int), ("transfers", C.c_void_p), ("hidden_cap", C.c_uint32),
                ("split_duration", C.c_float)]


class _Transfers(C.Structure):
    _fields_ = [("n", C.c_uint32), ("cap", C.c_uint32), ("source", C.c_void_p), ("target", C.c_void_p),
                ("time_left", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_kernel_height.restype = C.c_float
        _lib.orc_kernel_height.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float]
        _lib.orc_kernel_gradient.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        for f in ("orc_prefix_sum_helper_length", "orc_sort_helper_length"):
            getattr(_lib, f).restype = C.c_size_t
            getattr(_lib, f).argtypes = [C.c_size_t]
        for f in ("orc_neighborhood_green_pairs", "orc_neighborhood_binary_search_pairs",
                  "orc_neighborhood_brute_force_pairs", "orc_neighborhood_green_apply",
                  "orc_neighborhood_binary_search_apply", "orc_spread_kernel_width_apply", "orc_substep"):
            getattr(_lib, f).restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def default_settings():
    s = Settings()
    lib().orc_default_settings(C.byref(s))
    return s


def set_threads(n):
    lib().orc_set_threads(int(n))


class State:
    """The reference's particle lists (list_definitions.h:9-20) as numpy arrays, all-fluid scene."""
    FIELDS = (("index_list", np.uint32, 1), ("position", np.int32, 4), ("velocity", np.float32, 4),
              ("inverse_mass", np.float32, 1), ("radius", np.float32, 1), ("pos_backup", np.int32, 4),
              ("transferring", np.uint32, 1), ("target_radius", np.float32, 1), ("kernel_width", np.float32, 1),
              ("boundariness", np.float32, 1), ("boundary_distance", np.uint32, 1))

    def __init__(self, **arrays):
        for name, dt, w in self.FIELDS:
            a = np.ascontiguousarray(arrays[name], dtype=dt)
            setattr(self, name, a.reshape(-1, w) if w > 1 else a.reshape(-1))
        self.n_hidden = self.position.shape[0]
        self.n = self.index_list.shape[0]

    def copy(self):
        return State(**{name: getattr(self, name).copy() for name, _, _ in self.FIELDS})

    def c(self):
        st = _State()
        st.n_hidden, st.n = self.n_hidden, self.n
        for name, _, _ in self.FIELDS:
            setattr(st, name, getattr(self, name).ctypes.data)
        return st


class Transfers:
    """hidden_transfers (list_definitions.h:16-18): rows (source idx, target idx, time left), fixed capacity"""

    def __init__(self, cap, source=(), target=(), time_left=()):
        self.cap = int(cap)
        self.source = np.zeros(self.cap, np.uint32)
        self.target = np.zeros(self.cap, np.uint32)
        self.time_left = np.zeros(self.cap, np.float32)
        self.n = len(source)
        self.source[:self.n], self.target[:self.n], self.time_left[:self.n] = source, target, time_left

    def copy(self):
        return Transfers(self.cap, self.source[:self.n], self.target[:self.n], self.time_left[:self.n])

    def c(self):
        t = _Transfers()
        t.n, t.cap = self.n, self.cap
        t.source, t.target, t.time_left = self.source.ctypes.data, self.target.ctypes.data, self.time_left.ctypes.data
        return t

    def rows(self):
        return self.source[:self.n].copy(), self.target[:self.n].copy(), self.time_left[:self.n].copy()


def _with_room(st, cap):
    """the lists of `st` in arrays with room for `cap` entries (split appends particles)"""
    for name, dt, w in State.FIELDS:
        a = getattr(st, name)
        big = np.zeros((cap, w) if w > 1 else (cap,), dt)
        big[:a.shape[0]] = a
        setattr(st, name, big)


def _trim(st, cst):
    hidden = ("position", "velocity", "inverse_mass", "radius", "pos_backup", "transferring")
    for name, _, _ in State.FIELDS:
        setattr(st, name, getattr(st, name)[:cst.n_hidden if name in hidden else cst.n].copy())
    st.n, st.n_hidden = int(cst.n), int(cst.n_hidden)


# ---- thin wrappers --------------------------------------------------------------------
def kernel_height(s, dims, r, h):
    r = np.ascontiguousarray(r, dtype=np.float32)
    return lib().orc_kernel_height(C.byref(s), dims, _p(r), C.c_float(h))


def kernel_gradient(s, dims, r, h):
    r = np.ascontiguousarray(r, dtype=np.float32)
    o = np.zeros(3, np.float32)
    lib().orc_kernel_gradient(C.byref(s), dims, _p(r), C.c_float(h), _p(o))
    return o


def sort(keys, vals, upper_bound=0xFFFFFFFF):
    keys = np.ascontiguousarray(keys, np.uint32); vals = np.ascontiguousarray(vals, np.uint32)
    ok = np.empty_like(keys); ov = np.empty_like(vals)
    lib().orc_sort(_p(keys), _p(vals), C.c_uint32(len(keys)), C.c_uint32(upper_bound), _p(ok), _p(ov))
    return ok, ov


def prefix_sum(v):
    v = np.ascontiguousarray(v, np.uint32)
    o = np.empty_like(v)
    lib().orc_prefix_sum(_p(v), C.c_uint32(len(v)), _p(o))
    return o


def prefix_sum_helper_length(n):
    return lib().orc_prefix_sum_helper_length(n)


def sort_helper_length(n):
    return lib().orc_sort_helper_length(n)


def apply_edit(src, edit):
    src = np.ascontiguousarray(src)
    edit = np.ascontiguousarray(edit, np.uint32)
    stride = src.dtype.itemsize * (int(np.prod(src.shape[1:])) if src.ndim > 1 else 1)
    dst = np.empty((len(edit),) + src.shape[1:], src.dtype)
    lib().orc_apply_edit(_p(src), _p(dst), _p(edit), C.c_uint32(len(edit)), C.c_uint32(stride))
    return dst


def position_hash(pos4, mn, mx, res, dims):
    pos4 = np.ascontiguousarray(pos4, np.int32)
    out = np.empty(pos4.shape[0], np.uint32)
    lib().orc_position_hash(_p(pos4), C.c_uint32(pos4.shape[0]), _f3(mn), _f3(mx), C.c_uint32(res), dims, _p(out))
    return out


def position_code(index_list, pos4, section):
    index_list = np.ascontiguousarray(index_list, np.uint32)
    pos4 = np.ascontiguousarray(pos4, np.int32)
    out = np.empty(len(index_list), np.uint32)
    lib().orc_position_code(_p(index_list), _p(pos4), C.c_uint32(len(index_list)), C.c_uint32(section), _p(out))
    return out


def find_value_ranges(index_list, values, n_cells):
    index_list = np.ascontiguousarray(index_list, np.uint32)
    values = np.ascontiguousarray(values, np.uint32)
    s = np.zeros(n_cells, np.uint32); e = np.zeros(n_cells, np.uint32)
    lib().orc_find_value_ranges(_p(index_list), _p(values), C.c_uint32(len(index_list)), _p(s), _p(e))
    return s, e


def brute_force_pairs(index_list, pos4, rng, scale, cap):
    index_list = np.ascontiguousarray(index_list, np.uint32)
    pos4 = np.ascontiguousarray(pos4, np.int32); rng = np.ascontiguousarray(rng, np.float32)
    pairs = np.zeros((cap, 2), np.uint32)
    n = lib().orc_neighborhood_brute_force_pairs(_p(index_list), _p(pos4), _p(rng), C.c_uint32(len(index_list)),
                                                 C.c_float(scale), _p(pairs), C.c_uint32(cap))
    return pairs[:n]


def green_apply(st, s, dims, scale, mn, mx, res, cap, want_aux=False):
    """neighborhood_green::apply on `st` (modified in place). Returns pairs [P,2] (+aux dict)."""
    pairs = np.zeros((cap, 2), np.uint32)
    aux = {}
    if want_aux:
        aux = dict(sorted_hash=np.zeros(st.n_hidden, np.uint32), sorted_index=np.zeros(st.n_hidden, np.uint32),
                   cell_start=np.zeros(1 << (res * dims), np.uint32), cell_end=np.zeros(1 << (res * dims), np.uint32))
    cst = st.c()
    n = lib().orc_neighborhood_green_apply(C.byref(cst), C.byref(s), dims, C.c_float(scale), _f3(mn), _f3(mx),
                                           C.c_uint32(res), _p(pairs), C.c_uint32(cap),
                                           _p(aux.get("sorted_hash")), _p(aux.get("sorted_index")),
                                           _p(aux.get("cell_start")), _p(aux.get("cell_end")))
    return (pairs[:n], aux) if want_aux else pairs[:n]


def binary_search_apply(st, s, scale, cap, want_aux=False):
    pairs = np.zeros((cap, 2), np.uint32)
    aux = {}
    if want_aux:
        aux = dict(code0=np.zeros(st.n_hidden, np.uint32), code1=np.zeros(st.n_hidden, np.uint32),
                   code2=np.zeros(st.n_hidden, np.uint32), sorted_index=np.zeros(st.n_hidden, np.uint32))
    cst = st.c()
    n = lib().orc_neighborhood_binary_search_apply(C.byref(cst), C.byref(s), C.c_float(scale), _p(pairs), C.c_uint32(cap),
                                                   _p(aux.get("code0")), _p(aux.get("code1")), _p(aux.get("code2")),
                                                   _p(aux.get("sorted_index")))
    return (pairs[:n], aux) if want_aux else pairs[:n]


def incompressibility_apply(st, s, dims, pairs, want_aux=False):
    pairs = np.ascontiguousarray(pairs, np.uint32)
    incomp = np.zeros((st.n, 8), np.uint32) if want_aux else None
    lam = np.zeros(st.n, np.float32) if want_aux else None
    cst = st.c()
    lib().orc_incompressibility_apply(C.byref(cst), C.byref(s), dims, _p(pairs), C.c_uint32(len(pairs)), _p(incomp), _p(lam))
    if want_aux:
        return dict(grad_sum=incomp[:, 0:3].view(np.int32).copy(), density=incomp[:, 3].copy(),
                    sq_grad_sum=incomp[:, 4].copy(), lam=lam)


def spread_kernel_width_apply(st, s, pairs):
    """Returns (kept pairs, kw_fixed)."""
    pairs = np.ascontiguousarray(pairs, np.uint32).copy()
    kwfx = np.zeros(st.n, np.uint32)
    cst = st.c()
    n = lib().orc_spread_kernel_width_apply(C.byref(cst), C.byref(s), _p(pairs), C.c_uint32(len(pairs)), _p(kwfx))
    return pairs[:n], kwfx


def update_transfers_apply(st, s, pairs):
    """update_transfers::apply with merge and split off (update_transfers.cpp:14-54); returns the nearest-neighbour ids"""
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    nearest = np.zeros(st.n, np.uint32)
    cst = st.c()
    lib().orc_update_transfers_apply(C.byref(cst), C.byref(s), _p(pairs), C.c_uint32(len(pairs)), _p(nearest))
    return nearest


def update_transfers_full(st, s, dims, pairs, transfers, hidden_cap, split_duration=0.0):
    """update_transfers::apply (update_transfers.cpp:14-70) with merge / split as the settings say; `st` and `transfers`
    are updated in place (the lists may grow up to hidden_cap).  Returns the nearest-neighbour ids."""
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    nearest = np.zeros(max(st.n, 1), np.uint32)
    _with_room(st, hidden_cap)
    cst, ct = st.c(), transfers.c()
    lib().orc_update_transfers_full(C.byref(cst), C.c_uint32(hidden_cap), C.c_uint32(hidden_cap), C.byref(s), dims, _p(pairs),
                                    C.c_uint32(len(pairs)), C.byref(ct), C.c_float(split_duration), _p(nearest))
    _trim(st, cst)
    transfers.n = int(ct.n)
    return nearest


def particle_transfer_apply(st, transfers, dims, dt):
    """particle_transfer::apply(dt) (particle_transfer.cpp:10-28)"""
    cst, ct = st.c(), transfers.c()
    lib().orc_particle_transfer_apply(C.byref(cst), C.byref(ct), dims, C.c_float(dt))
    _trim(st, cst)
    transfers.n = int(ct.n)


def kernel_width_from_boundary_distance(st, s):
    """pool.cpp:77-80 (uint_to_float_with_indexed_lower_bound.comp)"""
    cst = st.c()
    lib().orc_kernel_width_from_boundary_distance(C.byref(cst), C.byref(s))


def box_collision(st, box_min4, box_max4):
    box_min4 = np.ascontiguousarray(box_min4, np.float32).reshape(-1, 4)
    box_max4 = np.ascontiguousarray(box_max4, np.float32).reshape(-1, 4)
    cst = st.c()
    lib().orc_box_collision(C.byref(cst), _p(box_min4), _p(box_max4), C.c_uint32(box_min4.shape[0]))


def velocity_handling(st, dt, accel):
    cst = st.c()
    lib().orc_velocity_handling(C.byref(cst), C.c_float(dt), _f3(accel))


def substep(st, s, *, dims, basic_pbf, solver_iterations, min_pos, max_pos, res_log2, box_min4, box_max4, cap,
            use_binary_search=False, integrate=False, dt=1.0 / 60.0, accel=(0.0, -10.0, 0.0), update_transfers=False,
            transfers=None, hidden_cap=None, split_duration=0.0):
    """One substep in pool::update order (pool.cpp:67-106). Returns the final pair list."""
    box_min4 = np.ascontiguousarray(box_min4, np.float32).reshape(-1, 4)
    box_max4 = np.ascontiguousarray(box_max4, np.float32).reshape(-1, 4)
    p = _SubstepParams()
    p.dims, p.basic_pbf, p.solver_iterations = dims, int(basic_pbf), solver_iterations
    p.use_binary_search, p.integrate, p.dt = int(use_binary_search), int(integrate), dt
    p.accel, p.min_pos, p.max_pos = _f3(accel), _f3(min_pos), _f3(max_pos)
    p.res_log2, p.n_boxes = res_log2, box_min4.shape[0]
    p.box_min4, p.box_max4 = box_min4.ctypes.data, box_max4.ctypes.data
    p.update_transfers = int(update_transfers)
    pairs = np.zeros((cap, 2), np.uint32)
    ct = None
    if transfers is not None:                       # settings::merge / split: the lists may grow up to hidden_cap
        hidden_cap = int(hidden_cap or st.n_hidden)
        _with_room(st, hidden_cap)
        ct = transfers.c()
        p.transfers, p.hidden_cap, p.split_duration = C.addressof(ct), hidden_cap, split_duration
    cst = st.c()
    n = lib().orc_substep(C.byref(cst), C.byref(s), C.byref(p), _p(pairs), C.c_uint32(cap))
    if transfers is not None:
        _trim(st, cst)
        transfers.n = int(ct.n)
    return pairs[:n]


# ---- the four incompressibility passes one by one (incompressibility.cpp:39-42), for callers that step between them ----
def incompressibility_passes_012(st, s, dims, pairs):
    """Passes 0, 1, 2 on `st` (positions get the particles' own shift, incompressibility_2.comp:106-109).
    Returns (incomp [n,8] u32, grad4 [P,4] f32, lambda [n] f32) for pass 3."""
    pairs = np.ascontiguousarray(pairs, np.uint32)
    n, npairs = st.n, len(pairs)
    incomp = np.zeros((max(n, 1), 8), np.uint32)
    lam = np.zeros(max(n, 1), np.float32)
    grad4 = np.zeros((max(npairs, 1), 4), np.float32)
    com4 = np.zeros((max(n, 1), 4), np.int32)
    cst = st.c()
    L = lib()
    L.orc_incompressibility_0(C.byref(cst), C.byref(s), dims, _p(incomp))
    L.orc_incompressibility_1(C.byref(cst), C.byref(s), dims, _p(pairs), C.c_uint32(npairs), _p(incomp), _p(com4), _p(grad4))
    L.orc_incompressibility_2(C.byref(cst), C.byref(s), dims, _p(incomp), _p(com4), _p(lam))
    return incomp, grad4, lam


def incompressibility_pass_3(st, s, dims, pairs, incomp, grad4, lam):
    pairs = np.ascontiguousarray(pairs, np.uint32)
    cst = st.c()
    lib().orc_incompressibility_3(C.byref(cst), C.byref(s), dims, _p(pairs), C.c_uint32(len(pairs)), _p(grad4), _p(lam), _p(incomp))
