"""ctypes binding of the C ABI in include/ragroute_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no CPU fallback: if the library is missing ``load()`` raises, and every
compute entry point fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libragroute_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ragroute_b200.h")

RR_MODE_AUTO = 0
RR_MODE_EXACT = 1
RR_MAX_K = 32
RR_SEARCH_NORMALIZE_L2 = 1
RR_MERGE_MAX_L = 4096

_lib = None


class RRError(RuntimeError):
    """Non-zero rr_status from the native library (faiss raises on bad input too)."""


class IndexStats(ctypes.Structure):
    _fields_ = [("searches", c_int64), ("queries", c_int64), ("exact_rescans", c_int64),
                ("scan_launches", c_int64), ("scan_ms_total", c_double), ("scan_bytes_total", c_int64),
                ("scan_flops_total", c_int64), ("exact_scans", c_int64)]


_P = c_void_p  # device / host pointers are passed as integers

_SIGNATURES = {
    "rr_abi_version": (c_int, []),
    "rr_last_error": (c_char_p, []),
    "rr_device_count": (c_int, []),
    "rr_kernel_launches": (c_int64, []),
    "rr_index_create": (c_int, [c_int, c_int64, c_int, POINTER(c_void_p)]),
    "rr_index_destroy": (c_int, [c_void_p]),
    "rr_index_reserve": (c_int, [c_void_p, c_int64]),
    "rr_index_add_f32": (c_int, [c_void_p, _P, c_int64, c_int, _P]),
    "rr_index_reset": (c_int, [c_void_p]),
    "rr_index_ntotal": (c_int64, [c_void_p]),
    "rr_index_dim": (c_int, [c_void_p]),
    "rr_index_reconstruct": (c_int, [c_void_p, c_int64, c_int64, _P, _P]),
    "rr_index_centroid": (c_int, [c_void_p, _P]),
    "rr_index_set_id_offset": (c_int, [c_void_p, c_int64]),
    "rr_index_set_param": (c_int, [c_void_p, c_char_p, c_double]),
    "rr_index_search": (c_int, [c_void_p, _P, c_int64, c_int, _P, _P, _P]),
    "rr_index_search_masked": (c_int, [c_void_p, _P, c_int64, c_int, _P, c_int64, _P, _P, _P]),
    "rr_index_search_ex": (c_int, [c_void_p, _P, c_int64, c_int, _P, c_int64, _P, _P, c_int64, _P]),
    "rr_index_search_host": (c_int, [c_void_p, _P, c_int64, c_int, _P, _P]),
    "rr_index_search_host_ex": (c_int, [c_void_p, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "rr_federation_search": (c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), c_int64, c_int, _P, c_int64, _P, _P,
                                     c_int64, _P]),
    "rr_index_read": (c_int, [c_char_p, c_int, POINTER(c_void_p)]),
    "rr_index_write": (c_int, [c_void_p, c_char_p]),
    "rr_index_get_stats": (c_int, [c_void_p, POINTER(IndexStats), c_int]),
    "rr_index_tc_qtile": (c_int, [c_void_p]),
    "rr_index_debug_tc_scan": (c_int, [c_void_p, _P, c_int, _P, _P, c_int64, POINTER(c_int), POINTER(c_int), _P]),
    "rr_normalize_l2": (c_int, [_P, c_int64, c_int, _P]),
    "rr_router_create": (c_int, [c_int, c_int, c_int] + [_P] * 10 + [c_int, POINTER(c_void_p)]),
    "rr_router_destroy": (c_int, [c_void_p]),
    "rr_router_logits": (c_int, [c_void_p, _P, c_int64, _P, _P]),
    "rr_router_set_tables": (c_int, [c_void_p, c_int, c_int, c_int, _P, _P, _P, c_int, _P, _P]),
    "rr_router_select": (c_int, [c_void_p, _P, c_int64, c_float, _P, _P, _P]),
    "rr_merge_topk": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "rr_merge_topk_segments": (c_int, [_P, _P, c_int64, c_int, c_int, c_int64, c_int64, c_int, c_int, c_int,
                                       _P, _P, _P, _P, _P, _P]),
}


def header_symbols() -> list:
    """Every function the public header declares (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the in-tree library and bind every entry point.  No CUDA call is made."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RRError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = symbol not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rr_abi_version() != 1:
        raise RRError("ABI version mismatch between _lib.py and libragroute_b200.so")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().rr_last_error().decode("utf-8", "replace")
        raise RRError(f"ragroute_b200 native error {rc}: {msg}")


def require_cuda() -> None:
    n = load().rr_device_count()
    if n <= 0:
        msg = load().rr_last_error().decode("utf-8", "replace")
        raise RRError(f"no CUDA device available ({msg}); ragroute_b200 has no CPU fallback")
