"""ctypes binding of libspeinet_b200.so (the C-ABI declared in include/speinet_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# SPEINET_B200_LIB selects an alternative build of the same library (A/B kernel experiments, tools/ab_*.sh);
# it is still this repo's CUDA library -- there is no other implementation to fall back to.
LIB_PATH = os.environ.get("SPEINET_B200_LIB") or os.path.join(_PKG, "libspeinet_b200.so")

FOLD_CUDA, FOLD_CPU = 0, 3
FOLD_ORDER_CPU, FOLD_TRUE_DIV = 1, 2
FOLD_LV1_PLANAR, FOLD_LV1_CELLS = 4, 8   # pin the source layout of the finest gather level (default: chosen on the device)
SEARCH_TC, SEARCH_EXACT, SEARCH_TCS = 0, 1, 2
IO_F32, IO_BF16 = 0, 1
STATS_WORDS = 8   # SPEI_STATS_WORDS
STATS_NAMES = ("saturated_queries", "pairs_rescored", "max_bf16_vs_exact_x1e9", "exhaustive_fallback", "reserved4",
               "second_pass_pairs_emitted", "certified_bound_violations", "reserved7")
VERSION = 200


class SpeiShape(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
        ("hr", ctypes.c_int32), ("wr", ctypes.c_int32), ("rf", ctypes.c_int32),
        ("c3", ctypes.c_int32), ("c2", ctypes.c_int32), ("c1", ctypes.c_int32),
        ("fold_mode", ctypes.c_int32), ("search", ctypes.c_int32), ("eps", ctypes.c_float),
        ("io_dtype", ctypes.c_int32),
    ]


_P = ctypes.c_void_p
_SH = ctypes.POINTER(SpeiShape)
# name -> (restype, argtypes); mirrors include/speinet_b200.h one to one
SIGNATURES = {
    "spei_version": (ctypes.c_int, []),
    "spei_last_error": (ctypes.c_char_p, []),
    "spei_workspace_bytes": (ctypes.c_int, [_SH, ctypes.POINTER(ctypes.c_size_t)]),
    "spei_search_transfer": (ctypes.c_int, [_SH, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "spei_stage_norm": (ctypes.c_int, [_SH, _P, _P, _P, ctypes.c_size_t, _P]),
    "spei_relevance_argmax": (ctypes.c_int, [_SH, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "spei_relevance_candidates": (ctypes.c_int, [_SH, _P, ctypes.c_size_t, _P]),
    "spei_rescore": (ctypes.c_int, [_SH, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "spei_gather_fold": (ctypes.c_int, [_SH, ctypes.c_int, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]),
    "spei_fuse_level": (ctypes.c_int, [ctypes.c_int32] * 5 + [_P] * 7),
    "spei_fuse_level_bf16": (ctypes.c_int, [ctypes.c_int32] * 5 + [_P] * 7),
    "spei_rl_deconv": (ctypes.c_int, [ctypes.c_int32] * 6 + [ctypes.c_float] + [_P] * 4),
    "spei_conv1x1": (ctypes.c_int, [ctypes.c_int32] * 3 + [ctypes.c_int64, _P, _P, _P, _P]),
    "spei_upsample2_bias_act": (ctypes.c_int, [ctypes.c_int32] * 4 + [_P, _P, ctypes.c_int32, _P, _P]),
    "spei_debug_relevance_tile": (ctypes.c_int, [_SH, _P, _P, ctypes.c_size_t, _P]),
    "spei_debug_error_flag": (ctypes.c_int, [_SH, _P, ctypes.c_size_t, _P, ctypes.POINTER(ctypes.c_int32)]),
    "spei_debug_search_cycles": (ctypes.c_int, [_SH, _P, ctypes.c_size_t, _P, ctypes.POINTER(ctypes.c_int64)]),
    "spei_plan_info": (ctypes.c_int, [_SH, ctypes.POINTER(ctypes.c_int32)]),
}

_lock = threading.Lock()
_lib = None


def load() -> ctypes.CDLL:
    """Load the library once.  Raises RuntimeError (never falls back) if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m speinet_b200.build` "
                    "(speinet_b200 has no CPU or PyTorch fallback)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here = header / library mismatch
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().spei_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
