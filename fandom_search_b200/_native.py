"""ctypes binding of the C ABI declared in include/fandom_search.h.

The library is built in-tree by ``fandom_search_b200.build`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing or no B200 is visible the search entry points raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfandom_search.so")

FS_OK = 0
FS_E_INVALID = -1
FS_E_CUDA = -2
FS_E_OVERFLOW = -3
FS_E_NOMEM = -4
FS_E_NODEVICE = -5

FS_CNT_CANDIDATES = 0
FS_CNT_MATCHES = 1
FS_CNT_EXACT = 2
FS_CNT_WINDOWS = 3
FS_CNT_OVERFLOW = 4
FS_CNT_ROWS = 5
FS_CNT_COUNT = 6
FS_OVERFLOW_CANDIDATES = 1
FS_OVERFLOW_MATCHES = 2
FS_OVERFLOW_ROWS = 4
FS_OVERFLOW_TEXT = 8

FS_OPT_SHIFTS_PER_STAGE = 1
FS_OPT_GRID_LIMIT = 3
FS_OPT_DIAG = 4
FS_OPT_CTA_PAIR = 5
FS_OPT_A_RESIDENT = 6
FS_OPT_PACKED_SHUFFLE = 7
FS_OPT_OPERAND_BITS = 9
FS_OPT_TILE_GROUP = 10
FS_OPT_PREFILTER_DIMS = 11
FS_OPT_FUSED_GATHER = 12

FS_MATCH_EXACT = 1
FS_MATCH_LSH_SHIFT = 8

# struct fs_match {int32 fan_pos; int32 script_pos; double distance; int32 work; uint32 flags;}
MATCH_DTYPE = np.dtype([("fan_pos", "<i4"), ("script_pos", "<i4"), ("distance", "<f8"),
                        ("work", "<i4"), ("flags", "<u4")], align=True)
PAIR_DTYPE = np.dtype([("fan_pos", "<i4"), ("script_pos", "<i4")], align=True)
# struct fs_row {int32 work, word, window_ix, match_ix; double distance; int32 lev, reserved;}
ROW_DTYPE = np.dtype([("work", "<i4"), ("word", "<i4"), ("window_ix", "<i4"), ("match_ix", "<i4"),
                      ("distance", "<f8"), ("lev", "<i4"), ("reserved", "<i4")], align=True)
assert MATCH_DTYPE.itemsize == 24 and PAIR_DTYPE.itemsize == 8 and ROW_DTYPE.itemsize == 32

# every symbol include/fandom_search.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_BATCH = [_vp, _i64, _vp, _i64]                 # tok, n_tok, off, n_works
_BATCHX = _BATCH + [_vp, _i64]                  # + extra, n_extra
SIGNATURES = {
    "fs_abi_version": (ctypes.c_int, []),
    "fs_last_error": (ctypes.c_char_p, []),
    "fs_device_count": (ctypes.c_int, []),
    "fs_index_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, _vp, _i64, _i32, _vp, _i64,
                                       _vp, _i64, _vp, _i64, _i32, _f64]),
    "fs_index_destroy": (ctypes.c_int, [_vp]),
    "fs_index_reserve": (ctypes.c_int, [_vp, _i64, _i64]),
    "fs_index_set_option": (ctypes.c_int, [_vp, _i32, _i64]),
    "fs_index_set_lsh": (ctypes.c_int, [_vp, _vp, _i32, _i32]),
    "fs_index_get_info": (_i64, [_vp, _i32]),
    "fs_search_csr_dev": (ctypes.c_int, [_vp, _vp] + _BATCHX + [_vp, _i64, _vp]),
    "fs_search_csr_host": (ctypes.c_int, [_vp] + _BATCHX + [_vp, _i64, _vp]),
    "fs_search_submit": (ctypes.c_int, [_vp] + _BATCHX + [_i64, ctypes.POINTER(_i32)]),
    "fs_search_collect": (ctypes.c_int, [_vp, _i32, _vp, _i64, _vp]),
    "fs_index_set_script_text": (ctypes.c_int, [_vp, ctypes.c_char_p, _vp, _i64]),
    "fs_search_submit_rows": (ctypes.c_int, [_vp] + _BATCHX + [_vp, _i64, _vp, _vp, _i32, _i64, _i64,
                                                               ctypes.POINTER(_i32)]),
    "fs_search_collect_rows": (ctypes.c_int, [_vp, _i32, _vp, _i64, _vp]),
    "fs_index_set_reuse_histogram": (ctypes.c_int, [_vp, _vp, _vp, _i32]),
    "fs_exact_join_dev": (ctypes.c_int, [_vp, _vp] + _BATCH + [_vp, _i64, _vp]),
    "fs_exact_join_host": (ctypes.c_int, [_vp] + _BATCH + [_vp, _i64, _vp]),
    "fs_stage_embed_dev": (ctypes.c_int, [_vp, _vp] + _BATCHX + [_vp, _vp]),
    "fs_stage_dots_dev": (ctypes.c_int, [_vp, _vp] + _BATCHX + [_vp, _i64]),
    "fs_stage_candidates_dev": (ctypes.c_int, [_vp, _vp] + _BATCHX + [_vp, _i64, _vp]),
    "fs_reuse_histogram_dev": (ctypes.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _i64, _vp]),
    "fs_timing_reset": (ctypes.c_int, [_vp]),
    "fs_timing_read": (ctypes.c_int, [_vp, ctypes.POINTER(_f64), ctypes.POINTER(_i64)]),
    "fs_index_scale": (ctypes.c_float, [_vp]),
    "fs_levenshtein_utf8": (_i32, [ctypes.c_char_p, _i64, ctypes.c_char_p, _i64]),
    "fs_murmurhash64a": (ctypes.c_uint64, [ctypes.c_char_p, _i64, ctypes.c_uint64]),
    "fs_tokenize_ws": (_i64, [ctypes.c_char_p, _i64, _vp, _vp, _i64]),
    "fs_vocab_create": (_vp, [ctypes.c_char_p, _vp, _vp, _i64]),
    "fs_vocab_destroy": (None, [_vp]),
    "fs_vocab_lookup": (_i32, [_vp, ctypes.c_char_p, _i64]),
    "fs_batch_encode_files": (_vp, [_vp, ctypes.POINTER(ctypes.c_char_p), _i64, _i32]),
    "fs_batch_destroy": (None, [_vp]),
    "fs_batch_info": (_i64, [_vp, _i32]),
    "fs_batch_array": (_vp, [_vp, _i32]),
    "fs_records_format_csv": (_i64, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_char_p, _vp, _vp, _vp, _vp,
                                     _vp, ctypes.c_char_p, _vp, _vp, ctypes.c_char_p, _vp, _vp, _vp, _vp, _i64,
                                     ctypes.POINTER(_vp)]),
    "fs_free": (None, [_vp]),
    "fs_format_py_float": (_i64, [_f64, ctypes.c_char_p, _i64]),
    "fs_records_best": (_i64, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64,
                               _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "fs_records_best_mt": (_i64, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32]),
}

_lib = None


class NativeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("fandom_search native error %d: %s" % (status, message))
        self.status = status


def load():
    """Load libfandom_search.so (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing: run `python -m fandom_search_b200.build` (or __graft_entry__.build()). "
            "This path has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fs_abi_version() != 2:
        raise RuntimeError("libfandom_search.so ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status != FS_OK:
        raise NativeError(status, load().fs_last_error().decode("utf-8", "replace"))
    return status


def ptr(a):
    """Address of a numpy array / torch tensor / None as a void pointer value."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor
