"""The C-ABI shared library loads and exports every symbol include/optcuts_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "optcuts_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header():
    from optcuts_b200 import build, _capi
    lib = build.build()                       # nvcc cross-compiles sm_100a without a GPU
    assert os.path.exists(lib)
    L = ctypes.CDLL(lib)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert set(names) == set(_capi.EXPORTED_SYMBOLS), set(names) ^ set(_capi.EXPORTED_SYMBOLS)


def test_header_is_plain_c(tmp_path):
    """include/optcuts_b200.h compiles as C99 (no C++/torch types in the signatures)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "optcuts_b200.h"\nint main(void){ocb_ctx* c=0; (void)c; return OCB_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_context_creation_is_lazy_and_errors_are_statuses():
    """ocb_create does no CUDA work (cheap shim objects); misuse returns status codes, never aborts."""
    import numpy as np
    from optcuts_b200 import _capi
    L = _capi.load_library()
    h = ctypes.c_void_p()
    assert L.ocb_create(ctypes.byref(h), 0) == 0
    assert L.ocb_launch_count(h) == 0
    s = np.zeros(8, np.int64)
    assert L.ocb_get_sizes(h, s.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))) == 0 and not s.any()
    assert L.ocb_set_uv(h, None, None) == -3           # OCB_ERR_STATE: no mesh yet
    assert b"ocb_set_mesh" in L.ocb_last_error(h)
    assert L.ocb_set_mesh(h, 0, 0, None, None, 1.0, None, 0) == -2      # OCB_ERR_ARG
    L.ocb_destroy(h)
    assert L.ocb_version().startswith(b"optcuts_b200")


def test_missing_library_fails_loudly(monkeypatch):
    from optcuts_b200 import _capi
    monkeypatch.setattr(_capi, "_LIB", None)
    monkeypatch.setenv("OPTCUTS_B200_LIB", "/nonexistent/liboptcuts_b200.so")
    with pytest.raises(_capi.OcbError):
        _capi.load_library()
