"""ctypes binding of ``libxspect_b200.so`` (C ABI: ``include/xspect_b200.h``).

The library is the only compute path.  If it is missing or cannot be loaded this module raises;
nothing here (or anywhere under ``xspect2_b200``) falls back to a CPU implementation.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libxspect_b200.so"

XS_OK = 0
XS_ERR_ARG, XS_ERR_IO, XS_ERR_FORMAT, XS_ERR_CUDA, XS_ERR_NOMEM, XS_ERR_UNSUPPORTED, XS_ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
XS_U8, XS_U16, XS_U32 = 1, 2, 4
XS_NONACGT_SKIP, XS_NONACGT_LITERAL = 0, 1
XS_COBS_CLASSIC, XS_COBS_COMPACT = 1, 2


class CobsInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("term_size", C.c_uint32), ("canonicalize", C.c_uint32), ("num_hashes", C.c_uint32),
        ("n_docs_total", C.c_uint32), ("doc_begin", C.c_uint32), ("doc_end", C.c_uint32), ("n_pages", C.c_uint32),
        ("page_bytes", C.c_uint64), ("row_stride", C.c_uint64), ("sig_size_max", C.c_uint64), ("hbm_bytes", C.c_uint64),
        ("device", C.c_int32), ("policy", C.c_int32),
    ]


class BloomInfo(C.Structure):
    _fields_ = [
        ("n_bits", C.c_uint64), ("k_hashes", C.c_uint64), ("term_size", C.c_uint32), ("device", C.c_int32),
        ("hbm_bytes", C.c_uint64),
    ]


class CobsHeader(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("term_size", C.c_uint32), ("canonicalize", C.c_uint32), ("num_hashes", C.c_uint32),
        ("n_docs", C.c_uint32), ("n_pages", C.c_uint32),
        ("page_bytes", C.c_uint64), ("sig_size_max", C.c_uint64), ("data_offset", C.c_uint64), ("file_size", C.c_uint64),
        ("layout", C.c_char * 64),
    ]


# every symbol include/xspect_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "xs_version": (C.c_int, []),
    "xs_last_error": (C.c_char_p, []),
    "xs_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "xs_launch_count": (C.c_uint64, []),
    "xs_device_trim": (C.c_int, [C.c_int]),
    "xs_profile_enable": (C.c_int, [C.c_int]),
    "xs_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "xs_profile_read_phases": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "xs_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(_P)]),
    "xs_host_free": (C.c_int, [_P]),
    "xs_cobs_open": (C.c_int, [C.c_char_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "xs_cobs_info": (C.c_int, [_P, C.POINTER(CobsInfo)]),
    "xs_cobs_doc_names": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "xs_cobs_header_layout": (C.c_char_p, [_P]),
    "xs_cobs_kernel": (C.c_char_p, [_P]),
    "xs_cobs_probe_header": (C.c_int, [C.c_char_p, C.POINTER(CobsHeader)]),
    "xs_cobs_doc_fill": (C.c_int, [_P, C.c_uint64, _P]),
    "xs_cobs_set_policy": (C.c_int, [_P, C.c_int]),
    "xs_cobs_set_bucketed": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32]),
    "xs_cobs_bucketed_queries": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "xs_bloom_set_bucketed": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int]),
    "xs_bloom_bucketed_queries": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "xs_cobs_close": (C.c_int, [_P]),
    "xs_cobs_query": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, C.c_int, _P]),
    "xs_cobs_query_device": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, C.c_int, _P, _P]),
    "xs_cobs_query_device_ld": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64, _P, _P]),
    "xs_comm_unique_id": (C.c_int, [_P]),
    "xs_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "xs_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xs_comm_destroy": (C.c_int, [_P]),
    "xs_allgather_scores": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P]),
    "xs_allreduce_totals": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "xs_sharded_reduce_device": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P]),
    "xs_cobs_create_synthetic": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(_P)]),
    "xs_cobs_classify": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, _P, _P, _P, _P]),
    "xs_mlst_query": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, _P, _P, _P]),
    "xs_cobs_result_order": (C.c_int, [_P, C.c_uint32, _P]),
    "xs_cobs_result_order_batch": (C.c_int, [_P, C.c_uint64, C.c_uint32, _P]),
    "xs_scores_reduce_device": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "xs_bloom_open": (C.c_int, [C.c_char_p, C.c_uint32, C.c_int, C.POINTER(_P)]),
    "xs_bloom_info": (C.c_int, [_P, C.POINTER(BloomInfo)]),
    "xs_bloom_close": (C.c_int, [_P]),
    "xs_bloom_query": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, _P]),
    "xs_bloom_contains": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "xs_bloom_query_device": (C.c_int, [_P, _P, C.c_uint64, _P, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "xs_cobs_build": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, C.c_uint64, C.c_uint64,
                                C.c_char_p, C.c_uint32, _P, C.c_uint64, _P, _P, _P, C.c_uint64]),
    "xs_bloom_build": (C.c_int, [C.c_char_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double, _P, C.c_uint64, _P, _P, C.c_uint64]),
    "xs_fastx_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "xs_fastx_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "xs_fastx_read": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "xs_fastx_filter_fasta": (C.c_int, [_P, _P, C.c_char_p]),
    "xs_fastx_close": (C.c_int, [_P]),
    "xs_cobs_classify_file": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(_P)]),
    "xs_file_calls_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "xs_file_calls_read": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "xs_file_calls_view": (C.c_int, [_P] + [C.POINTER(_P)] * 7),
    "xs_file_calls_free": (C.c_int, [_P]),
    "xs_result_write_json": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, _P, C.c_uint32, _P, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "xs_pack_2bit": (C.c_int, [_P, C.c_uint64, C.c_int, _P, _P]),
    "xs_canonical_kmers": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_int, _P, _P]),
    "xs_cobs_rows": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "xs_kmer_rows": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, _P, _P]),
    "xs_bloom_hashes": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P]),
}

_lib = None


class XsError(RuntimeError):
    """A C-ABI call failed; ``code`` is the negative xs_status."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def lib_path() -> Path:
    return Path(os.environ.get("XSPECT_B200_LIB", _LIB_PATH))


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise ImportError(
                f"{p} is missing: build it with `make` (nvcc, sm_100a). xspect2_b200 has no CPU fallback."
            )
        L = C.CDLL(str(p))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Map an xs_status to the exception type the reference raises in the same situation."""
    if rc == XS_OK:
        return
    msg = lib().xs_last_error().decode("utf-8", "replace")
    if rc == XS_ERR_IO:
        raise FileNotFoundError(msg)
    if rc in (XS_ERR_ARG, XS_ERR_FORMAT):
        raise ValueError(msg)
    if rc == XS_ERR_NOMEM:
        raise MemoryError(msg)
    raise XsError(rc, msg)
