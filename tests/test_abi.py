"""CPU: the C-ABI library loads without a GPU and exports every symbol that
include/solidboolean_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

import solidboolean_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "solidboolean_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(sb.LIB_PATH):
        from solidboolean_b200.build import build
        build()
    return C.CDLL(sb.LIB_PATH)


def test_header_declares_functions():
    names = declared_functions()
    assert len(names) >= 30
    assert "sb_mesh_create" in names and "sb_intersect" in names and "sb_classify" in names


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_table_matches_header():
    assert sorted(sb.ABI) == declared_functions()


def test_header_cites_reference_for_each_group():
    src = open(HEADER).read()
    for cite in ("src/solidmesh.cpp:42-76", "src/solidboolean.cpp:94-101", "src/solidboolean.cpp:48-92",
                 "tri_tri_intersect.c:395-472", "axisalignedboundingboxtree.h:54-95"):
        assert cite in src


def test_no_device_fails_loudly(lib):
    """Without a GPU the product path must refuse to run, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib.sb_context_create.restype = C.c_int
    lib.sb_last_error.restype = C.c_char_p
    h = C.c_void_p()
    rc = lib.sb_context_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.sb_last_error() or b"CUDA" in lib.sb_last_error()
    with pytest.raises(sb.SolidBooleanError):
        sb.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under solidboolean_b200/ may use it."""
    pkg = os.path.join(ROOT, "solidboolean_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "sbo_" not in text and "libsboracle" not in text and "libsbref" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
