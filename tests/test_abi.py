"""The C-ABI library loads without a GPU and exports every symbol include/gpb200.h declares; the ctypes
signature table mirrors the header.  No compute calls here."""
import ctypes
import os
import re

from conftest import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int|long|size_t|void|const char\s*\*)\s+(gpb_\w+)\s*\(([^;{]*)\)\s*;", text)
    return {name: args for name, args in decls}


def test_library_exports_every_declared_symbol():
    from gptorch_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    fns = _header_functions()
    assert len(fns) >= 20
    for name in fns:
        assert hasattr(lib, name), "libgpb200.so does not export %s" % name


def test_ctypes_table_matches_header():
    from gptorch_b200 import _lib
    fns = _header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, args in fns.items():
        args = args.strip()
        n_header = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n_header == len(_lib.SIGNATURES[name][1]), name


def test_load_and_metadata():
    from gptorch_b200 import _lib
    lib = _lib.load()
    assert lib.gpb_version() >= 100
    assert lib.gpb_block_size() == 128
    assert lib.gpb_last_error() is not None
    # size queries are pure host functions
    assert lib.gpb_potri_workspace_bytes(32768) >= 32768 * 16384 * 8 // 2
    assert lib.gpb_trsv_workspace_bytes(1000) >= 8 * 4
    assert lib.gpb_kern_bwd_workspace_bytes(1000, 1000, 8) > 0
    assert lib.gpb_gpr_grad_workspace_bytes(1000, 8) > 0


def test_header_cites_reference_call_sites():
    text = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    for cite in ("gptorch/kernels.py", "gptorch/functions.py:46-47", "gptorch/functions.py:71-76",
                 "gptorch/functions.py:61-68", "gptorch/models/gpr.py", "gptorch/util.py:73-88"):
        assert cite in text


def test_plain_c_client_compiles_links_and_validates_arguments(tmp_path):
    """include/gpb200.h is a C header (C99, -pedantic) and the library is usable from plain C: a C program resolves
    every entry point with dlsym and gets the documented status codes from argument validation -- no GPU needed."""
    import subprocess
    from gptorch_b200 import _lib
    src = os.path.join(ROOT, "tests", "abi", "abi_check.c")
    exe = str(tmp_path / "abi_check")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    src, "-o", exe, "-ldl"], check=True)
    out = subprocess.run([exe, _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    fns = _header_functions()
    assert "abi_check: %d symbols ok" % len(fns) in out, out
    # the C program's symbol list is the header's
    listed = set(re.findall(r'"(gpb_\w+)"', open(src).read()))
    assert listed == set(fns)
