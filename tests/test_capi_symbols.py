"""The C-ABI library loads without a GPU and exports every symbol include/rxmesh_b200.h declares."""
import ctypes
import os
import re

import rxmesh_b200 as rx
from rxmesh_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_declared_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "rxmesh_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rxm_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    so = ctypes.CDLL(rx.LIB_PATH)
    for n in sorted(names):
        assert hasattr(so, n), n
    assert names == set(_lib.SYMBOLS), names ^ set(_lib.SYMBOLS)


def test_version_and_error_string():
    assert b"sm_100a" in rx.lib().rxm_version()
    assert rx.lib().rxm_mesh_info(None, 0) == 0


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/rxmesh_b200.h must compile as C99 (no torch / C++ types in the signatures)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        return
    src = tmp_path / "c_abi.c"
    src.write_text('#include "rxmesh_b200.h"\nint main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "c_abi.o")])
