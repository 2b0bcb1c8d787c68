"""ctypes binding of libvqb200.so (include/vqb200.h).

The library is the product: if it is missing or no sm_100 GPU is present, everything
here raises -- there is no CPU path and nothing in this package imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VQB200_LIB: a differently built library (A/B timings of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("VQB200_LIB") or os.path.join(HERE, "libvqb200.so")

SUCCESS, ERR_NULL_PTR, ERR_EMPTY_INPUT, ERR_INVALID_INPUT = 0, -1, -2, -3
ERR_UNSUPPORTED_DEVICE, ERR_DIM_MISMATCH, FAILURE = -4, -5, -99

METRIC_IDS = {"squared_euclidean": 0, "euclidean": 1, "manhattan": 2, "cosine": 3,
              "chebyshev": 5}   # 5 = VQB_CHEBYSHEV: extension, not in the reference
UPDATE_ORDERED, UPDATE_FAST = 0, 1
ASSIGN_AUTO, ASSIGN_EXACT, ASSIGN_TENSOR = 0, 1, 2
TRAIN_USE_COMM = 1
COMM_ID_BYTES = 128

RESEED_FN = C.CFUNCTYPE(C.c_uint64, C.c_void_p, C.c_uint32)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class TrainOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("update_mode", C.c_uint32), ("assign_mode", C.c_uint32),
        ("flags", C.c_uint32),
        ("reseed", RESEED_FN), ("reseed_user", C.c_void_p),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p),
        ("row_offset", C.c_uint64), ("n_global", C.c_uint64),
        ("iter_ms", C.POINTER(C.c_float)),
    ]


# name -> (restype, argtypes); mirrors include/vqb200.h one to one
_P = C.c_void_p
_SZ = C.c_size_t
SIGNATURES = {
    "vqb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vqb_ctx_destroy": (C.c_int, [_P]),
    "vqb_ctx_synchronize": (C.c_int, [_P]),
    "vqb_ctx_stream": (_P, [_P]),
    "vqb_ctx_set_stream": (C.c_int, [_P, _P]),
    "vqb_last_error": (C.c_char_p, [_P]),
    "vqb_backend_name": (C.c_char_p, []),
    "vqb_ctx_launch_count": (C.c_uint64, [_P]),
    "vqb_malloc": (C.c_int, [_P, _SZ, C.POINTER(_P)]),
    "vqb_free": (C.c_int, [_P, _P]),
    "vqb_host_alloc": (C.c_int, [_P, _SZ, C.POINTER(_P)]),
    "vqb_host_free": (C.c_int, [_P, _P]),
    "vqb_memcpy": (C.c_int, [_P, _P, _P, _SZ]),
    "vqb_comm_unique_id": (C.c_int, [_P]),
    "vqb_comm_init_rank": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "vqb_comm_destroy": (C.c_int, [_P]),
    "vqb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vqb_comm_allreduce": (C.c_int, [_P, _P, _SZ]),
    "vqb_distance_batch": (C.c_int, [_P, C.c_int, _P, _P, _SZ, _SZ, _P]),
    "vqb_bq_quantize": (C.c_int, [_P, _P, _SZ, C.c_float, C.c_uint8, C.c_uint8, _P]),
    "vqb_bq_dequantize": (C.c_int, [_P, _P, _SZ, C.c_uint8, C.c_uint8, _P]),
    "vqb_sq_quantize": (C.c_int, [_P, _P, _SZ, C.c_float, C.c_float, C.c_float, C.c_uint32, _P]),
    "vqb_sq_dequantize": (C.c_int, [_P, _P, _SZ, C.c_float, C.c_float, _P]),
    "vqb_f16_dequantize": (C.c_int, [_P, _P, _SZ, _P]),
    "vqb_pq_train": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, _SZ, _SZ, _P, C.POINTER(TrainOpts), _P, _P]),
    "vqb_pq_assign_train": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, _SZ, _P, C.c_uint32, _P]),
    "vqb_pq_train_step": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, _SZ, _P, C.POINTER(TrainOpts), _P, _P]),
    "vqb_pq_create": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, C.c_int, C.POINTER(_P)]),
    "vqb_pq_destroy": (C.c_int, [_P]),
    "vqb_pq_codebooks": (C.c_int, [_P, _P]),
    "vqb_pq_encode": (C.c_int, [_P, _P, _SZ, C.c_uint32, _P, C.c_uint32, _P]),
    "vqb_pq_decode": (C.c_int, [_P, _P, C.c_uint32, _SZ, _P]),
    "vqb_debug_tc_scores": (C.c_int, [_P, C.c_int, _P, _SZ, _SZ, _SZ, _SZ, _P, C.c_int, _P, _P, _P]),
    "vqb_debug_tc_timeline": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, _SZ, _P, _P, C.c_int]),
    "vqb_debug_tc_variant": (C.c_int, [C.c_int]),
    "vqb_tsvq_train": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, C.c_int, C.POINTER(_P)]),
    "vqb_tsvq_create": (C.c_int, [_P, _P, _P, _P, _SZ, _SZ, C.c_int, C.POINTER(_P)]),
    "vqb_tsvq_destroy": (C.c_int, [_P]),
    "vqb_tsvq_num_nodes": (C.c_int, [_P, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "vqb_tsvq_export": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "vqb_tsvq_encode": (C.c_int, [_P, _P, _SZ, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libvqb200.so; raises if it has not been built (python -m vq_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m vq_b200.build` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
