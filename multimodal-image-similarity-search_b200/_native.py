"""ctypes binding of ``libvecsearch_b200.so`` (C ABI: ``include/vecsearch_b200.h``).

There is no fallback of any kind: if the shared library is missing or a call fails, an
exception is raised.  ``load()`` never builds anything -- ``build.py`` / ``__graft_entry__.build()``
does that.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libvecsearch_b200.so"
# VS_LIB_PATH: load another build of the same library (profiling experiments only)
LIB_PATH = os.environ.get("VS_LIB_PATH") or os.path.join(_HERE, LIB_NAME)

VS_F32, VS_BF16 = 0, 1
VS_Q_AUTO, VS_Q_SCAN, VS_Q_TENSOR = 0, 1, 2
VS_ERR_OVERFLOW = -5
VS_ERR_EXCHANGE = -6
VS_Q_PIPELINED = 0x100
MASK_WORDS = 4
MAX_K = 1024

# every symbol include/vecsearch_b200.h declares: name -> (restype, argtypes)
_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_f = C.c_float
SYMBOLS = {
    "vs_last_error": (C.c_char_p, []),
    "vs_abi_version": (_i, []),
    "vs_create": (_i, [_i, _i, _i, _i64, C.POINTER(_p)]),
    "vs_destroy": (_i, [_p]),
    "vs_count": (_i64, [_p]),
    "vs_dim": (_i, [_p]),
    "vs_dtype": (_i, [_p]),
    "vs_device": (_i, [_p]),
    "vs_reserve": (_i, [_p, _i64]),
    "vs_add_raw_host": (_i, [_p, _p, _i64, C.POINTER(_i64)]),
    "vs_get_raw_host": (_i, [_p, _i64, _i64, _p]),
    "vs_remove_rows": (_i, [_p, _p, _i64, _p, _p, C.POINTER(_i64)]),
    "vs_truncate": (_i, [_p, _i64]),
    "vs_copy_row": (_i, [_p, _i64, _p, _i64]),
    "vs_move_rows": (_i, [_p, _p, _p, _p, _i64]),
    "vs_replicate_from": (_i, [_p, _p, _i64, _i64]),
    "vs_get_mask_bits_range": (_i, [_p, _i64, _i64, _p]),
    "vs_apply_sweep_bits_dev": (_i, [_p, _p, _i, _p]),
    "vs_exchange_clear_error": (_i, [_p]),
    "vs_group_create": (_i, [_i, _p, _i, _i, _i64, _i, _i, C.POINTER(_p)]),
    "vs_group_destroy": (_i, [_p]),
    "vs_group_size": (_i, [_p]),
    "vs_group_shard": (_p, [_p, _i]),
    "vs_group_count": (_i64, [_p]),
    "vs_group_last_timing": (_i, [_p, _p]),
    "vs_group_query_host": (_i, [_p, _p, _i, _i, _p, _i, _p, _p]),
    "vs_group_query_multimodal_host": (_i, [_p, _p, _p, _p, _i, _i, _p, _i, _p, _p]),
    "vs_set_row_base": (_i, [_p, _i64]),
    "vs_set_row_map": (_i, [_p, _i64, _i64]),
    "vs_add_host": (_i, [_p, _p, _i64, C.POINTER(_i64)]),
    "vs_add_dev": (_i, [_p, _p, _i64, C.POINTER(_i64), _p]),
    "vs_remove": (_i, [_p, _i64, C.POINTER(_i64)]),
    "vs_set_row_host": (_i, [_p, _i64, _p]),
    "vs_clear": (_i, [_p]),
    "vs_set_mask_bits": (_i, [_p, _i64, _p]),
    "vs_get_mask_bits": (_i, [_p, _i64, _p]),
    "vs_set_mask_bits_range": (_i, [_p, _i64, _i64, _p]),
    "vs_get_rows_host": (_i, [_p, _i64, _i64, _p]),
    "vs_get_rows_dev": (_i, [_p, _i64, _i64, _p, _p]),
    "vs_query_topk_host": (_i, [_p, _p, _i, _i, _p, _i, _p, _p]),
    "vs_query_topk_dev": (_i, [_p, _p, _i, _i, _p, _i, _p, _p, _p]),
    "vs_blend_dev": (_i, [_p, _p, _p, _p, _i, _p, _p]),
    "vs_query_multimodal_host": (_i, [_p, _p, _p, _p, _i, _i, _p, _i, _p, _p]),
    "vs_merge_topk_dev": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "vs_exchange_bytes": (C.c_size_t, [_i, _i, _i]),
    "vs_exchange_create": (_i, [_p, _i, _i, _i, _i]),
    "vs_exchange_ipc_handle": (_i, [_p, _p]),
    "vs_exchange_local_ptr": (_p, [_p]),
    "vs_exchange_attach": (_i, [_p, _p, _p]),
    "vs_query_topk_sharded_dev": (_i, [_p, _p, _i, _i, _p, _i, _p, _p, _p]),
    "vs_exchange_merge_dev": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "vs_exchange_error": (_i, [_p]),
    "vs_query_topk_sharded_host": (_i, [_p, _p, _i, _i, _p, _i, _p, _p]),
    "vs_exchange_begin": (_i, [_p]),
    "vs_query_topk_push_dev": (_i, [_p, _p, _i, _i, _p, _i, _i, _p]),
    "vs_exchange_collect_dev": (_i, [_p, _i, _i, _p, _p, _p]),
    "vs_filter_words": (_i64, [_p]),
    "vs_filter_sweep_dev": (_i, [_p, _p, _i, _f, _p, _p]),
    "vs_filter_sweep_host": (_i, [_p, _p, _i, _f, _p]),
    "vs_dedup_dev": (_i, [_p, _i64, _i64, _f, _i64, _p, _p, _p, _p, _p]),
    "vs_dedup_host": (_i, [_p, _i64, _i64, _f, _i64, _p, _p, _p, C.POINTER(_i64)]),
    "vs_launch_count": (C.c_uint64, []),
    "vs_last_query_path": (_i, [_p]),
    "vs_device_sm_count": (_i, [_p]),
}


class VecSearchError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vecsearch_b200 error {code}: {msg}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the C-ABI library (once).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  This engine has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        raise VecSearchError(rc, load().vs_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().vs_launch_count())
