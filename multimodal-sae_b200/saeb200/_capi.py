"""ctypes binding of `lib/libsaeb200.so` (C ABI declared in include/saeb200.h).

The library is the product: there is no Python / PyTorch fallback.  If the shared object is missing the import of
this module raises; if a call fails the library's own error string is raised as `SaebError`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint32, c_void_p

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libsaeb200.so")

F32, BF16, F16 = 0, 1, 2


class SaebError(RuntimeError):
    pass


# name -> (restype, argtypes); must list every symbol include/saeb200.h declares (tests/test_capi_symbols.py)
SIGNATURES = {
    "saeb_version": (c_int, []),
    "saeb_last_error": (c_char_p, []),
    "saeb_launch_count": (c_longlong, []),
    "saeb_set_option": (c_int, [c_char_p, c_int]),
    "saeb_profile_last_encode_ms": (c_float, []),
    "saeb_debug_stats": (c_int, [c_void_p]),
    "saeb_query": (c_longlong, [c_char_p]),
    "saeb_packed_weights_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "saeb_packed_bias_offset": (c_size_t, [c_int64, c_int64, c_int]),
    "saeb_pack_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "saeb_encode_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "saeb_encode_topk": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64, c_int,
                                 c_int64, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t,
                                 c_void_p]),
    "saeb_encode_topk_refine_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "saeb_encode_topk_refine": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                        c_int, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_int, c_void_p]),
    "saeb_prep_bytes": (c_size_t, [c_int64, c_int64]),
    "saeb_prep_activations": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "saeb_candidates_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "saeb_encode_candidates": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int,
                                       c_int64, c_float, c_void_p, c_size_t, c_void_p]),
    "saeb_refine_candidates": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                       c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_float, c_void_p, c_void_p,
                                       c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_int, c_int, c_void_p]),
    "saeb_refine_candidates_lo": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                          c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_float, c_void_p,
                                          c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_size_t, c_int, c_int, c_void_p]),
    "saeb_candidate_bounds": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64, c_int,
                                      c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "saeb_candidate_bounds_packed": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64,
                                             c_int, c_int, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "saeb_dense_topk": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "saeb_decode": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int64, c_int64, c_void_p,
                            c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "saeb_decode_backward_acts": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64,
                                          c_void_p, c_void_p, c_void_p]),
    "saeb_decode_backward_weight": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64,
                                            c_void_p, c_void_p, c_void_p]),
    "saeb_coo_window_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p,
                                       c_void_p]),
    "saeb_column_sums": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "saeb_feature_maps": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "saeb_total_variance": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "saeb_coo_workspace_bytes": (c_size_t, [c_int64]),
    "saeb_coo_extract": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "saeb_coo_append": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_int64, c_int64, c_void_p,
                                c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "saeb_scan_pool": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_int64, c_int64, c_int64,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "saeb_scan_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "saeb_image_pool_workspace_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "saeb_image_pool_init": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int64, c_void_p]),
    "saeb_image_pool": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_int64, c_int64, c_int64,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "saeb_scan_pool_workspace_bytes": (c_size_t, [c_int, c_int]),
    "saeb_scan_pool_init": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "saeb_scan_pool_ws": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_int64, c_int64, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "saeb_debug_coload": (c_int, [c_int, c_int, c_int64, c_void_p, c_size_t, c_void_p, c_void_p]),
    "saeb_gathered_bounds": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "saeb_kth_of_gathered": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "saeb_kth_largest_gathered": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "saeb_push_gather": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_int, c_size_t, c_void_p, c_size_t, c_int,
                                 c_uint32, c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built: run `python __graft_entry__.py`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SaebError(
                f"{LIB_PATH} not found: the CUDA extension is not built (run `python __graft_entry__.py build`). "
                "There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().saeb_last_error().decode("utf-8", "replace")
        raise SaebError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().saeb_launch_count())
