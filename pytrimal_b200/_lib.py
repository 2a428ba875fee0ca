"""ctypes binding of libtrimal_cuda.so (the C ABI in include/trimal_cuda.h).

The library is the product; there is no Python or CPU implementation behind
it.  Importing this module without the built library raises ImportError, and
every compute call without a B200 raises :class:`NoDeviceError`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtrimal_cuda.so")

TCU_OK = 0
TCU_ERR_NO_DEVICE = -1
TCU_ERR_OOM = -2
TCU_ERR_CUDA = -3
TCU_ERR_INVALID = -4
TCU_ERR_INCORRECT_SYMBOL = -5
TCU_ERR_UNDEFINED_SYMBOL = -6
TCU_ERR_STATE = -7
TCU_ERR_NCCL = -8


class TrimalCudaError(RuntimeError):
    """Any failure reported by libtrimal_cuda (mapped to RuntimeError like
    pytrimal maps trimAl's generic error codes, src/trimal/source/reportsystem.cpp:43-52)."""

    def __init__(self, code, message):
        super().__init__(f"[trimal_cuda {code}] {message}")
        self.code = code


class NoDeviceError(TrimalCudaError):
    pass


class SymbolError(ValueError):
    """IncorrectSymbol / UndefinedSymbol (ValueError in pytrimal, reportsystem.cpp:43-52)."""

    def __init__(self, code, message, col, row, byte):
        super().__init__(message)
        self.code, self.col, self.row, self.byte = code, col, row, byte


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("kernel_ms", C.c_float),
                ("d2h_ms", C.c_float), ("kernel_launches", C.c_int), ("comm_ms", C.c_float)]


_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)
_h = C.c_void_p

#: name -> (restype, argtypes); must list every symbol include/trimal_cuda.h declares
PROTOTYPES = {
    "tcu_device_count": (C.c_int, []),
    "tcu_last_error": (C.c_char_p, []),
    "tcu_version": (C.c_char_p, []),
    "tcu_msa_create": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.POINTER(_h)]),
    "tcu_msa_create_strided": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                         C.POINTER(_h)]),
    "tcu_msa_destroy": (None, [_h]),
    "tcu_msa_create_all": (C.c_int, [_h, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(_h)]),
    "tcu_set_devices": (C.c_int, [_i32p, C.c_int]),
    "tcu_get_devices": (C.c_int, [_i32p, C.c_int]),
    "tcu_msa_device_count": (C.c_int, [_h]),
    "tcu_release_cached_memory": (None, []),
    "tcu_msa_nseq": (C.c_int, [_h]),
    "tcu_msa_ncol": (C.c_int, [_h]),
    "tcu_gaps": (C.c_int, [_h, _i32p, _i32p, _i32p, _i32p]),
    "tcu_identity": (C.c_int, [_h, _i32p, _i32p, C.c_uint8, _f32p, _i32p, _i32p, C.c_int]),
    "tcu_similarity": (C.c_int, [_h, C.c_uint8, _f32p, C.c_int, _i32p, _i32p, C.c_float, _f32p,
                                 _f32p, _f32p, _f32p, _i32p, _i32p, _i32p]),
    "tcu_spurious": (C.c_int, [_h, C.c_uint8, C.c_uint32, _f32p]),
    "tcu_identity_band": (C.c_int, [_h, _i32p, _i32p, C.c_uint8, C.c_int, C.c_int, _f32p]),
    "tcu_identity_band_rows": (C.c_int, []),
    "tcu_identity_row_blocks": (C.c_int, [C.c_int]),
    "tcu_identity_tile": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_longlong, _i32p, _i32p]),
    "tcu_identity_tiles_before": (C.c_longlong, [C.c_int, C.c_int]),
    "tcu_identity_row_offset": (C.c_size_t, [C.c_int, C.c_int]),
    "tcu_identity_prepare": (C.c_int, [_h, _i32p, _i32p, C.c_uint8, _i32p]),
    "tcu_identity_device": (C.c_int, [_h, C.c_int, C.c_int, C.c_void_p]),
    "tcu_msa_sync": (C.c_int, [_h]),
    "tcu_msa_stream": (C.c_void_p, [_h]),
    "tcu_msa_device": (C.c_int, [_h]),
    "tcu_msa_timings": (C.c_int, [_h, C.POINTER(Timings)]),
    "tcu_comm_id": (C.c_int, [C.c_void_p]),
    "tcu_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(_h)]),
    "tcu_comm_destroy": (None, [_h]),
    "tcu_comm_rank": (C.c_int, [_h]),
    "tcu_comm_world": (C.c_int, [_h]),
    "tcu_shard_range": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
    "tcu_shard_blocks": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
    "tcu_identity_all": (C.c_int, [_h, _h, _i32p, _i32p, C.c_uint8, _f32p]),
    "tcu_similarity_all": (C.c_int, [_h, _h, C.c_uint8, _f32p, C.c_int, _i32p, _i32p, C.c_float,
                                     _f32p, _f32p, _f32p, _i32p, _i32p, _i32p]),
    "tcu_gaps_all": (C.c_int, [_h, _h, _i32p, _i32p, _i32p, _i32p]),
    "tcu_spurious_all": (C.c_int, [_h, _h, C.c_uint8, C.c_uint32, _f32p]),
    "tcu_identity_resident": (C.c_int, [_h]),
    "tcu_identity_download": (C.c_int, [_h, _f32p]),
    "tcu_identity_row_stats": (C.c_int, [_h, C.c_int, _f32p, _f32p, _f32p]),
    "tcu_identity_clusters": (C.c_int, [_h, _i32p, C.c_int, C.c_float, _i32p, _i32p]),
    "tcu_byte_histogram": (C.c_int, [_h, C.POINTER(C.c_ulonglong)]),
    "tcu_sequence_lengths": (C.c_int, [_h, _i32p]),
    "tcu_row_residues": (C.c_int, [_h, _i32p, _i32p]),
    "tcu_row_hashes": (C.c_int, [_h, C.POINTER(C.c_ulonglong)]),
    "tcu_cluster_order": (C.c_int, [_i32p, C.c_int, _i32p]),
    "tcu_threshold_rule": (None, [C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_int)]),
    "tcu_representatives": (C.c_int, [_h, _i32p, C.c_uint8, C.c_float, _i32p, _i32p]),
    "tcu_representatives_all": (C.c_int, [_h, _h, _i32p, C.c_uint8, C.c_float, _i32p, _i32p]),
}

_lib = None


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m pytrimal_b200.build` "
                "(nvcc, sm_100a).  pytrimal_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().tcu_last_error().decode(errors="replace")


def check(rc, err=None):
    if rc == TCU_OK:
        return
    msg = last_error()
    if rc in (TCU_ERR_INCORRECT_SYMBOL, TCU_ERR_UNDEFINED_SYMBOL):
        col, row, byte = err if err else (-1, -1, 0)
        raise SymbolError(rc, msg, col, row, byte)
    if rc == TCU_ERR_NO_DEVICE:
        raise NoDeviceError(rc, msg)
    if rc == TCU_ERR_OOM:
        raise MemoryError(msg)
    if rc == TCU_ERR_INVALID:
        raise ValueError(msg)
    raise TrimalCudaError(rc, msg)


def device_count() -> int:
    return load().tcu_device_count()
