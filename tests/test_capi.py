"""The C-ABI library loads and exports exactly what include/pcc_b200.h declares (no compute: runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcc_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from pcc_b200 import _lib
    return _lib


def test_header_declares_entry_points():
    names = _declared()
    for required in ("pcc_fps_f32", "pcc_knn_f32", "pcc_ball_query_f32", "pcc_gather_f32", "pcc_gather_bwd_f32",
                     "pcc_nn1_f32", "pcc_chamfer_fwd_f32", "pcc_chamfer_bwd_f32", "pcc_version",
                     "pcc_last_error_string"):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (pcc_[a-z0-9_]+)", out))
    declared = set(_declared())
    assert declared <= exported, f"declared but not exported: {declared - exported}"
    assert exported <= declared, f"exported but not declared in the header: {exported - declared}"
    assert set(lib.SIGNATURES) == declared
    assert L.pcc_version() >= 100


def test_sass_is_sm100a_only(lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", lib.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_do_not_need_a_gpu(lib):
    L = lib.load()
    rc = L.pcc_knn_f32(None, None, 1, 1, 1, 1, None, None, None, 0, 1.0, None)
    assert rc == -1 and b"null pointer" in L.pcc_last_error_string()
    rc = L.pcc_knn_f32(8, 8, 1, 4, 4, 5000, 8, 8, None, 0, 1.0, None)
    assert rc == -1 and b"K=5000" in L.pcc_last_error_string()
