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


def test_exchange_layout_and_peer_entry_points_validate_without_a_gpu(lib):
    """The query-sliced exchange lays its per-rank buffer out on the host (pure arithmetic): flags, then two parities
    of {gather scores, gather ids, result scores, result ids}; arguments are validated before any CUDA call."""
    need = ctypes.c_size_t(0)
    assert lib.drs_exchange_sliced_bytes(65536, 65536 * 100 + 800, 8, ctypes.byref(need)) == 0
    entries = 65536 * 100 + 800
    assert need.value >= 2 * 24 * entries and need.value % 256 == 0          # 2 parities x (4 + 8 + 4 + 8) bytes per entry
    small = ctypes.c_size_t(0)
    assert lib.drs_exchange_sliced_bytes(1000, 10000, 2, ctypes.byref(small)) == 0 and small.value < need.value
    assert lib.drs_exchange_sliced_bytes(0, 10, 2, ctypes.byref(need)) == 1
    assert lib.drs_exchange_sliced_bytes(10, 10, 9, ctypes.byref(need)) == 1   # at most 8 peers
    assert lib.drs_exchange_sliced_bytes(10, 10, 2, None) == 1
    assert lib.drs_exchange_sliced(None, None, 4, 5, 0, 2, None, 4, 100, None, None, None, None) == 1
    assert lib.drs_peer_alloc(0, None, None) == 1 and lib.drs_peer_open(None, None) == 1
    assert lib.drs_peer_close(None) == 0 and lib.drs_peer_free(None) == 0      # nothing to release
    _lib.set_option("tune.cooperative", 0)
    assert _lib.get_option("tune.cooperative") == 0
    _lib.set_option("tune.cooperative", 1)
    assert _lib.get_option("search.fp32_mode") == 0 and _lib.get_option("tune.symmetric_grad") == 2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib._build, "LIB", str(tmp_path / "libdrs_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
