"""ctypes binding of the C ABI declared in include/drs_b200.h.

There is no CPU or torch fallback: if the CUDA extension is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

DRS_F32, DRS_BF16, DRS_F16 = 0, 1, 2
DRS_MAX_K = 256

_lock = threading.Lock()
_lib = None

c_i64, c_int, c_f32, c_vp, c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

_SIGNATURES = {
    "drs_version": (c_int, []),
    "drs_last_error": (ctypes.c_char_p, []),
    "drs_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "drs_get_option": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_int)]),
    "drs_debug_max_clusters": (c_int, [c_int, ctypes.POINTER(c_int)]),
    "drs_debug_triangle_walk": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int)]),
    "drs_debug_hang_report": (c_int, [ctypes.POINTER(ctypes.c_uint * 6)]),
    "drs_search_workspace_bytes": (c_int, [c_i64, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_sz)]),
    "drs_search": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_debug_open_claims": (c_int, [c_vp, ctypes.POINTER(ctypes.c_uint * 8), c_vp]),
    "drs_search_scan": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_sz, c_vp]),
    "drs_search_select": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "drs_search_l2_workspace_bytes": (c_int, [c_i64, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_sz)]),
    "drs_search_l2": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_rerank": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "drs_pair_scores": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "drs_cluster_update": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "drs_doc_pairs_workspace_bytes": (c_int, [c_i64, c_i64, ctypes.POINTER(c_sz)]),
    "drs_doc_sentence_pairs": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp,
                                       c_sz, c_vp]),
    "drs_exchange_flag_bytes": (c_int, [c_i64, c_int, ctypes.POINTER(c_sz)]),
    "drs_search_sharded_p2p": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp,
                                       c_sz, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_exchange_sliced_bytes": (c_int, [c_i64, c_i64, c_int, ctypes.POINTER(c_sz)]),
    "drs_exchange_sliced": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "drs_peer_alloc": (c_int, [c_sz, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_ubyte * 64)]),
    "drs_peer_open": (c_int, [ctypes.POINTER(ctypes.c_ubyte * 64), ctypes.POINTER(c_vp)]),
    "drs_peer_close": (c_int, [c_vp]),
    "drs_peer_free": (c_int, [c_vp]),
    "drs_merge_shards": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp]),
    "drs_infonce_workspace_bytes": (c_int, [c_i64, c_int, c_i64, c_int, ctypes.POINTER(c_sz)]),
    "drs_infonce_forward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_f32, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_infonce_backward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_f32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_sz, c_vp]),
    "drs_infonce_backward_staged": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_f32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                            c_sz, c_vp]),
    "drs_moco_workspace_bytes": (c_int, [c_i64, c_int, c_i64, c_int, ctypes.POINTER(c_sz)]),
    "drs_moco_forward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_f32, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_moco_backward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_f32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  c_sz, c_vp]),
    "drs_proto_workspace_bytes": (c_int, [c_i64, c_int, c_i64, c_int, ctypes.POINTER(c_sz)]),
    "drs_proto_forward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "drs_proto_backward": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_sz,
                                   c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libdrs_b200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"drs_b200: CUDA extension not built ({path} missing). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root; "
                "there is no CPU fallback.")
        lib = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    """C status -> Python exception, message from drs_last_error().  The reference's callers
    catch RuntimeError (src/train.py:190-195), so that is what is raised."""
    if rc != 0:
        msg = load().drs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"drs_b200 error {rc}: {msg}")


_options_epoch = 0


def options_epoch() -> int:
    """bumped by every set_option: callers that cache layout answers (workspace sizes) key them on it"""
    return _options_epoch


def set_option(name: str, value: int):
    global _options_epoch
    _options_epoch += 1
    check(load().drs_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = c_int(0)
    check(load().drs_get_option(name.encode(), ctypes.byref(v)))
    return v.value


def hang_report():
    """{flag, tag, block, thread, parity, extra} of the last timed-out pipeline wait (debug)."""
    out = (ctypes.c_uint * 6)()
    check(load().drs_debug_hang_report(ctypes.byref(out)))
    return dict(zip(("flag", "tag", "block", "thread", "parity", "extra"), list(out)))
