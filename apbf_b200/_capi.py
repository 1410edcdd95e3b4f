"""ctypes declarations for libapbf_b200.so (include/apbf_b200.h).  No fallback: if the library is missing or has
no CUDA device to run on, callers get an exception."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libapbf_b200.so")

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class Settings(C.Structure):
    """apbf_settings (shaders/cpu_gpu_shared_config.h:74-94)"""
    _fields_ = [(n, C.c_int) for n in (
        "mHeightKernelId", "mGradientKernelId", "mMerge", "mSplit", "mBaseKernelWidthOnTargetRadius",
        "mBaseKernelWidthOnBoundaryDistance", "mUpdateTargetRadius", "mUpdateBoundariness",
        "mNeighborListSorted", "mBoundarinessCalculationMethod")] + [(n, C.c_float) for n in (
        "mBoundarinessAdaptionSpeed", "mKernelWidthAdaptionSpeed", "mBoundarinessSelfGradLengthFactor",
        "mBoundarinessUnderpressureFactor", "mMergeDuration", "mSmallestTargetRadius", "mTargetRadiusOffset",
        "mTargetRadiusScaleFactor")]


class Array(C.Structure):
    _fields_ = [("data", vp), ("reorder_out", vp)]


class Particles(C.Structure):
    _fields_ = [("index_list", Array), ("length", vp), ("capacity", C.c_uint32), ("hidden_length", vp),
                ("hidden_capacity", C.c_uint32), ("position", Array), ("velocity", Array), ("inverse_mass", Array),
                ("radius", Array), ("pos_backup", Array), ("transferring", Array)]


class Fluid(C.Structure):
    _fields_ = [("particle", Particles), ("target_radius", Array), ("kernel_width", Array), ("boundariness", Array),
                ("boundary_distance", Array)]


class Neighbors(C.Structure):
    _fields_ = [("pairs", vp), ("length", vp), ("capacity", C.c_uint32)]


class SearchDebug(C.Structure):
    _fields_ = [("sorted_key", vp), ("sorted_index", vp), ("cell_start", vp), ("cell_end", vp), ("code", vp * 3), ("pair_offsets", vp)]


class SimConfig(C.Structure):
    _fields_ = [("particle_capacity", C.c_uint32), ("neighbor_capacity", C.c_uint32), ("dims", C.c_int),
                ("basic_pbf", C.c_int), ("solver_iterations", C.c_int), ("use_binary_search", C.c_int),
                ("integrate", C.c_int), ("dt", C.c_float), ("accel", C.c_float * 3), ("min_pos", C.c_float * 3),
                ("max_pos", C.c_float * 3), ("res_log2", C.c_uint32), ("n_boxes", C.c_uint32), ("box_min4_host", vp),
                ("box_max4_host", vp), ("update_transfers", C.c_int), ("transfers", C.c_int),
                ("transfer_capacity", C.c_uint32), ("split_duration", C.c_float)]


class Transfers(C.Structure):
    _fields_ = [("source", Array), ("target", Array), ("time_left", Array), ("length", vp), ("capacity", C.c_uint32)]


class HostState(C.Structure):
    _fields_ = [("n", C.c_uint32)] + [(n, vp) for n in (
        "index_list", "position", "velocity", "inverse_mass", "radius", "pos_backup", "transferring", "target_radius",
        "kernel_width", "boundariness", "boundary_distance")]


# name -> (restype, argtypes); every symbol include/apbf_b200.h declares
SIGNATURES = {
    "apbf_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "apbf_ctx_destroy": (None, [vp]),
    "apbf_ctx_set_stream": (C.c_int, [vp, vp]),
    "apbf_ctx_synchronize": (C.c_int, [vp]),
    "apbf_ctx_last_error": (C.c_char_p, [vp]),
    "apbf_default_settings": (None, [C.POINTER(Settings)]),
    "apbf_ctx_set_settings": (C.c_int, [vp, C.POINTER(Settings)]),
    "apbf_ctx_set_dimensions": (C.c_int, [vp, C.c_int]),
    "apbf_ctx_launch_count": (C.c_uint64, [vp]),
    "apbf_ctx_set_search_stats": (C.c_int, [vp, C.c_int]),
    "apbf_ctx_set_stream_blocks": (C.c_int, [vp, C.c_uint32]),
    "apbf_ctx_set_match_grid_min": (C.c_int, [vp, C.c_uint32]),
    "apbf_ctx_device_flags": (C.c_int, [vp, u32p]),
    "apbf_ctx_list_state": (C.c_int, [vp, u32p]),
    "apbf_ctx_profile": (C.c_int, [vp, C.c_int]),
    "apbf_ctx_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "apbf_buffer_acquire": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "apbf_buffer_release": (C.c_int, [vp, vp]),
    "apbf_copy_bytes": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "apbf_copy_bytes_from_host": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "apbf_copy_bytes_to_host": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "apbf_write_sequence": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "apbf_write_sequence_float": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_float, C.c_float]),
    "apbf_copy_scattered_read": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_scattered_write": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_append_list": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "apbf_copy_with_differing_stride": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "apbf_find_value_changes": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32]),
    "apbf_write_increasing_sequence": (C.c_int, [vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_write_increasing_sequence_from_to": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32]),
    "apbf_apply_hidden_edit": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]),
    "apbf_sort": (C.c_int, [vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32]),
    "apbf_prefix_sum": (C.c_int, [vp, vp, vp, C.c_uint32, vp]),
    "apbf_sort_calculate_needed_helper_list_length": (C.c_size_t, [C.c_size_t]),
    "apbf_prefix_sum_calculate_needed_helper_list_length": (C.c_size_t, [C.c_size_t]),
    "apbf_calculate_position_hash": (C.c_int, [vp, vp, vp, vp, C.c_uint32, f32p, f32p, C.c_uint32]),
    "apbf_calculate_position_code": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_find_value_ranges": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_neighborhood_green_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Array), C.POINTER(Neighbors), C.c_float,
                                                f32p, f32p, C.c_uint32, C.POINTER(SearchDebug)]),
    "apbf_neighborhood_green_spread_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), C.c_float, f32p, f32p,
                                                       C.c_uint32, C.POINTER(SearchDebug), vp]),
    "apbf_neighborhood_binary_search_spread_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), C.c_float, vp, vp]),
    "apbf_neighborhood_binary_search_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Array), C.POINTER(Neighbors),
                                                        C.c_float, C.POINTER(SearchDebug)]),
    "apbf_neighbors_invalidate": (C.c_int, [vp, C.POINTER(Neighbors)]),
    "apbf_neighbors_release": (C.c_int, [vp, C.POINTER(Neighbors)]),
    "apbf_neighbors_prepare": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), vp]),
    "apbf_incompressibility_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), vp, vp]),
    "apbf_spread_kernel_width_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), vp]),
    "apbf_update_transfers_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), vp]),
    "apbf_update_transfers_split_merge_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Neighbors), C.POINTER(Transfers), C.c_float, vp]),
    "apbf_particle_transfer_apply": (C.c_int, [vp, C.POINTER(Fluid), C.POINTER(Transfers), C.c_float, vp]),
    "apbf_transfers_follow_reorder": (C.c_int, [vp, C.POINTER(Transfers), vp, vp, C.c_uint32]),
    "apbf_kernel_width_from_boundary_distance": (C.c_int, [vp, C.POINTER(Fluid)]),
    "apbf_uint_to_float_with_indexed_lower_bound": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_float, C.c_float, C.c_float]),
    "apbf_box_collision_apply": (C.c_int, [vp, C.POINTER(Particles), vp, vp, C.c_uint32]),
    "apbf_velocity_handling_apply": (C.c_int, [vp, C.POINTER(Particles), C.c_float, C.c_float, f32p]),
    "apbf_sim_create": (C.c_int, [vp, C.POINTER(SimConfig), C.POINTER(vp)]),
    "apbf_sim_destroy": (None, [vp]),
    "apbf_sim_upload": (C.c_int, [vp, C.POINTER(HostState)]),
    "apbf_sim_download": (C.c_int, [vp, C.POINTER(HostState)]),
    "apbf_sim_substep": (C.c_int, [vp, C.c_uint32]),
    "apbf_sim_fluid": (C.c_int, [vp, C.POINTER(Fluid)]),
    "apbf_sim_neighbors": (C.c_int, [vp, C.POINTER(Neighbors)]),
    "apbf_sim_neighbor_count": (C.c_int, [vp, u32p]),
    "apbf_sim_transfers": (C.c_int, [vp, C.POINTER(Transfers)]),
    "apbf_sim_download_transfers": (C.c_int, [vp, u32p, vp, vp, vp]),
    "apbf_sim_stats": (C.c_int, [vp, u32p]),
    "apbf_sim_mg_enable": (C.c_int, [vp, C.c_int, C.c_int, C.c_float]),
    "apbf_sim_mg_brick": (C.c_int, [vp, C.c_int, u32p, u32p, u32p]),
    "apbf_sim_mg_set_counts": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "apbf_sim_mg_route": (C.c_int, [vp, vp]),
    "apbf_sim_mg_pack_state": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "apbf_sim_mg_unpack_state": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, C.c_int]),
    "apbf_sim_mg_copy_state": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "apbf_sim_mg_swap": (C.c_int, [vp]),
    "apbf_sim_mg_halo_lists": (C.c_int, [vp, vp, C.c_uint32, vp]),
    "apbf_sim_mg_pack": (C.c_int, [vp, C.c_int, vp, C.c_uint32, vp]),
    "apbf_sim_mg_unpack": (C.c_int, [vp, C.c_int, vp, C.c_uint32, C.c_uint32, vp]),
    "apbf_sim_mg_remap": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32]),
    "apbf_sim_mg_phase": (C.c_int, [vp, C.c_int, C.c_int]),
    "apbf_mg_nccl_unique_id": (C.c_int, [vp]),
    "apbf_sim_mg_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "apbf_sim_mg_solve": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int]),
    "apbf_sim_mg_halo_counts": (C.c_int, [vp, C.c_uint32, u32p]),
    "apbf_sim_mg_loop_init": (C.c_int, [vp, C.c_uint32, C.c_uint32, u32p]),
    "apbf_sim_mg_loop_reset": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
    "apbf_sim_mg_loop_stats": (C.c_int, [vp, u32p]),
    "apbf_sim_mg_substep": (C.c_int, [vp, C.c_uint32]),
    "apbf_sim_step_host": (C.c_int, [vp, C.POINTER(HostState), C.POINTER(HostState)]),
    "apbf_sim_set_graphs": (C.c_int, [vp, C.c_int]),
    "apbf_sim_graph_replays": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "apbf_sim_mg_p2p_export": (C.c_int, [vp, vp]),
    "apbf_sim_mg_p2p_import": (C.c_int, [vp, vp]),
    "apbf_sim_mg_p2p_active": (C.c_int, [vp]),
    "apbf_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "apbf_host_free_pinned": (C.c_int, [vp]),
}

_lib = None


def load():
    """dlopen the library and declare every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m apbf_b200.build` "
                               "(there is no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
