"""The C++ operator surface (include/apbf_pbd.hpp: pbd::gpu_list / indexed_list / uninterleaved_list and the operator
classes) over the C-ABI.  CPU leg: the reference's list tests restated in tests/cpp/test_pbd_lists.cpp compile and link
against libapbf_b200.so.  GPU leg: they run (source/test.cpp:47-531 known-answer vectors + one search/solve)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_pbd_lists.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_pbd_lists")


def _build():
    from apbf_b200 import build
    build.build()
    deps = [SRC, os.path.join(ROOT, "include", "apbf_pbd.hpp"), os.path.join(ROOT, "include", "apbf_b200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(d) <= os.path.getmtime(EXE) for d in deps):
        return EXE
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + os.path.join(ROOT, "apbf_b200"),
           "-lapbf_b200", "-Wl,-rpath," + os.path.join(ROOT, "apbf_b200"), "-Wl,-rpath,$ORIGIN/../../apbf_b200", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_cpp_shim_compiles_and_links():
    exe = _build()
    assert os.path.exists(exe)
    # without a device the binary must refuse to run (exit code 77), never compute on the host
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_shim_reference_list_tests():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "all pbd list / algorithm / operator tests passed" in r.stdout
