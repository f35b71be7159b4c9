"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and
exports every symbol include/ftb200.h declares; the ctypes table covers them all.
No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from femtech_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ftb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ftb200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ["ftb200_create", "ftb200_upload_mesh", "ftb200_shape_functions", "ftb200_lumped_mass",
                 "ftb200_get_force", "ftb200_calculate_accelerations", "ftb200_stable_time_step",
                 "ftb200_check_energy", "ftb200_explicit_begin", "ftb200_explicit_run", "ftb200_halo_pack",
                 "ftb200_halo_add"]:
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(lib_path):
    L = C.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(L, s), "libftb200.so does not export %s" % s


def test_ctypes_table_matches_header(lib_path):
    from femtech_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "ftb200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "#include" not in code
    assert 'extern "C"' in code


def test_create_fails_loudly_without_gpu(lib_path):
    """No CPU fallback: on a box without a CUDA device the product refuses to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from femtech_b200 import mesh, solver
    X, conn, pid = mesh.cube_mesh(2)
    with pytest.raises(solver.FemTechB200Error):
        solver.FemTech(X, conn, pid, [1], [1040.0, 100.0, 100.0, 0, 0, 0, 0, 0, 0])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure; nothing under femtech_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "femtech_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "femtech_oracle" not in txt and "oracle/" not in txt, f


def test_entry_points_reject_a_null_context(lib_path):
    """Argument checks of the C-ABI run before any CUDA call: a NULL context is refused with the reference's bad-input
    code (3) -- or the documented sentinel -- not a crash.  (No GPU needed; nothing is computed.)"""
    from femtech_b200 import _lib
    L = _lib.load()
    ring = C.POINTER(C.c_double)()
    assert L.ftb200_step_ring(None, 16, C.byref(ring)) == 3
    assert L.ftb200_affine_element_count(None) == -1
    assert L.ftb200_explicit_run_async(None, 1.0, 10) == 3
    assert L.ftb200_explicit_poll(None, None, None, None, None) == 3
    assert L.ftb200_shape_functions(None, None) == 3
    assert L.ftb200_launch_count(None) == 0
    assert b"null context" in L.ftb200_last_error(None)
