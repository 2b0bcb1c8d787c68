"""vq_b200 -- B200-native engine for the vq hot path (PQ/TSVQ training + encoding, BQ/SQ).

Public surface mirrors the reference's quantizer API (see vq_b200/api.py); the compute lives
in libvqb200.so (hand-written sm_100a CUDA behind the C ABI of include/vqb200.h).
"""
from .api import (BinaryQuantizer, DimensionMismatch, Distance, EmptyInput, Engine, FfiError,
                  InvalidParameter, ProductQuantizer, ScalarQuantizer, TSVQ, VqError, default_engine,
                  draw_init_indices, get_simd_backend)

__all__ = ["BinaryQuantizer", "ScalarQuantizer", "ProductQuantizer", "TSVQ", "Distance", "Engine",
           "VqError", "DimensionMismatch", "EmptyInput", "InvalidParameter", "FfiError",
           "default_engine", "draw_init_indices", "get_simd_backend"]
__version__ = "0.1.0"
