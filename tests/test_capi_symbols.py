"""CPU test: the C-ABI library builds, loads, and exports every symbol include/apbf_b200.h declares.
No compute call is made here (there is no GPU on the CPU test box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "apbf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apbf_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from apbf_b200 import build
    return ctypes.CDLL(build.build())


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 40
    for must in ("apbf_neighborhood_green_apply", "apbf_neighborhood_binary_search_apply", "apbf_incompressibility_apply",
                 "apbf_spread_kernel_width_apply", "apbf_sort", "apbf_prefix_sum", "apbf_sim_substep"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol():
    from apbf_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _declared()
    _capi.load()  # declares argtypes for every symbol; raises on a missing one


def test_helper_lengths_match_reference_formulas(lib):  # algorithms.cpp:38-58, host-only entry points
    lib.apbf_prefix_sum_calculate_needed_helper_list_length.restype = ctypes.c_size_t
    lib.apbf_sort_calculate_needed_helper_list_length.restype = ctypes.c_size_t
    f = lib.apbf_prefix_sum_calculate_needed_helper_list_length
    assert f(ctypes.c_size_t(512 * 512 + 1000)) == 514 + 2 + 1 + 10
    assert f(ctypes.c_size_t(7)) == 11
    assert lib.apbf_sort_calculate_needed_helper_list_length(ctypes.c_size_t(8)) == 16 + 11


def test_no_device_means_loud_failure(lib):
    """without a CUDA device context creation fails with APBF_ERR_NO_DEVICE; with one it succeeds"""
    h = ctypes.c_void_p()
    rc = lib.apbf_ctx_create(0, None, ctypes.byref(h))
    import torch
    if torch.cuda.is_available():
        assert rc == 0
        lib.apbf_ctx_destroy(h)
    else:
        assert rc == -1 and not h.value


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under apbf_b200/ or include/ may reference it"""
    bad = []
    for base in ("apbf_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"apbf_oracle|from oracle|import oracle|orc_[a-z]", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
