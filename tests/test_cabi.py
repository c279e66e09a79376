"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/rj_b200.h
declares, the Python binding covers exactly those symbols, and -- without a GPU -- the engine fails
loudly instead of falling back to a CPU path.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import helpers as H
from radix_join_b200 import _cabi

HEADER = os.path.join(H.ROOT, "include", "rj_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rj_[a-z0-9_]+)\s*\(", src)))


def _build_engine():
    if not os.path.exists(_cabi.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(H.ROOT, "radix-join_b200", "csrc"), "-j", "8"],
                       check=True, capture_output=True)


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(_cabi.PROTOTYPES) == names


def test_library_exports_every_declared_symbol():
    _build_engine()
    lib = C.CDLL(_cabi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/rj_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(rj_[a-z0-9_]+)", out))
    assert exported == set(declared_functions())


def test_load_library_binds_all_prototypes():
    _build_engine()
    lib = _cabi.load_library()
    assert lib.rj_version().startswith(b"radix-join_b200")
    assert lib.rj_fixed_rows_per_page(0) == 1984 and lib.rj_fixed_rows_per_page(1) == 1007
    assert [lib.rj_stage_name(i).decode() for i in range(_cabi.RJ_ST_COUNT)] == _cabi.STAGE_NAMES


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.EngineMissing):
        _cabi.load_library(str(tmp_path / "nope.so"))


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _build_engine()
    import radix_join_b200 as rj
    with pytest.raises(rj.EngineError) as e:
        rj.build_context()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under radix-join_b200/ may reference it"""
    pkg = os.path.join(H.ROOT, "radix-join_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "rj_oracle" not in text and "libref_oracle" not in text, f


def test_header_is_plain_c_and_the_c_example_binds(tmp_path):
    """The boundary is a C ABI: the header must compile as C11 (no C++), and a C caller written against it
    (examples/c_abi_join.c) must only need symbols the library exports.  Compile only -- no GPU here."""
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                   check=True, capture_output=True)
    obj = str(tmp_path / "c_abi_join.o")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.dirname(HEADER), "-c",
                    os.path.join(H.ROOT, "examples", "c_abi_join.c"), "-o", obj], check=True, capture_output=True)
    wanted = {line.split()[-1] for line in subprocess.run(["nm", obj], capture_output=True, text=True).stdout.splitlines()
              if " U rj_" in line}
    assert {"rj_ctx_create", "rj_execute", "rj_execute_streamed", "rj_result_fetch", "rj_ctx_destroy"} <= wanted
    _build_engine()
    exported = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    for name in wanted:
        assert re.search(rf"\bT {name}\b", exported), f"{name} is used by the C example but not exported"
