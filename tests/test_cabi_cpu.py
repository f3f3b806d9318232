"""The C-ABI library loads and exports every symbol include/*.h declares; argument validation
that needs no GPU.  CPU only: no compute calls."""
import ctypes
import glob
import os
import re

import pytest

import drs_b200
from drs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    drs_b200.build.build()            # no-op when libdrs_b200.so is newer than its sources
    return _lib.load()


def _declared_symbols():
    names = []
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libdrs_b200.so"
    # and the ctypes binding covers every declared symbol
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_string(lib):
    assert lib.drs_version() >= 100
    assert isinstance(lib.drs_last_error(), bytes)


def test_options_roundtrip_and_unknown_option(lib):
    _lib.set_option("search.cta_group", 1)
    assert _lib.get_option("search.cta_group") == 1
    _lib.set_option("search.cta_group", 0)
    with pytest.raises(RuntimeError, match="unknown option"):
        _lib.set_option("no.such.option", 1)


def test_null_and_range_checks_return_codes(lib):
    need = ctypes.c_size_t(0)
    # argument errors are reported before any CUDA call
    assert lib.drs_search_workspace_bytes(0, 10, 64, 5, _lib.DRS_BF16, ctypes.byref(need)) == 1
    assert b"positive" in lib.drs_last_error()
    assert lib.drs_search_workspace_bytes(4, 10, 64, 0, _lib.DRS_BF16, ctypes.byref(need)) == 3
    assert lib.drs_search_workspace_bytes(4, 10, 64, _lib.DRS_MAX_K + 1, _lib.DRS_BF16, ctypes.byref(need)) == 3
    assert lib.drs_search_workspace_bytes(4, 10, 64, 5, _lib.DRS_BF16, None) == 1
    assert lib.drs_search(None, 4, None, 10, 64, _lib.DRS_BF16, 5, 0, None, None, None, 0, None) == 1
    assert lib.drs_merge_shards(None, None, 2, 4, 5, None, None, None) == 1
    assert lib.drs_infonce_workspace_bytes(0, 64, 0, _lib.DRS_F32, ctypes.byref(need)) == 1
    assert lib.drs_infonce_forward(None, None, None, 4, 64, 0, 20.0, 0, None, None, None, 0, None) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib._build, "LIB", str(tmp_path / "libdrs_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
