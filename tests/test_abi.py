"""The C-ABI library builds, loads and exports every symbol include/mw_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mw_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_three_seams():
    syms = header_symbols()
    for s in ("mw_logmel", "mw_logmel_long", "mw_encode", "mw_generate", "mw_model_create", "mw_model_destroy",
              "mw_last_error", "mw_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(built_lib)], text=True)
    exported = set(re.findall(r" T (mw_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in include/mw_b200.h but not exported: {missing}"


def test_ctypes_signatures_cover_header(built_lib):
    from manual_whisper_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.mw_abi_version() == 2
    assert lib.mw_storage_dtype() in (0, 1)
    assert isinstance(lib.mw_launch_count(), int)


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from manual_whisper_b200 import _lib
    lib = _lib.load()
    handle = C.c_void_p()
    st = lib.mw_logmel_plan_create(128, None, 4, 0, C.byref(handle))
    assert st == 1 and b"null" in lib.mw_last_error()
    with pytest.raises(ValueError):
        _lib.check(st, "mw_logmel_plan_create")
    assert lib.mw_model_create(None, None, C.byref(handle)) == 1
    assert lib.mw_logmel(None, None, 0, None, None, 1, None, None, None) == 1


def test_sass_is_blackwell_native(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG must be present in the sm_100a SASS."""
    sass = subprocess.check_output(["cuobjdump", "-sass", str(built_lib)], text=True)
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, f"{mnemonic} missing: the GEMM/attention kernels are not on the tcgen05/TMA path"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "manual_whisper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
