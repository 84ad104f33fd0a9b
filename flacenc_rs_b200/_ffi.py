"""ctypes view of the C ABI in include/flacenc_b200.h and loader of libflacenc_b200.so.

There is no CPU fallback: `lib()` raises if the CUDA library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), and every encode call fails with
FB200_ERR_CUDA when no device is usable."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libflacenc_b200.so")

OK, ERR_CONFIG, ERR_SOURCE, ERR_CUDA, ERR_CAPACITY = 0, 1, 2, 3, 4
MAX_CHANNELS = 8
MAX_LPC_ORDER = 24
MAX_RICE_PARTS = 256
SF_CONSTANT, SF_VERBATIM, SF_FIXED, SF_LPC = 0, 1, 2, 3


class Config(C.Structure):
    """fb200_config: POD mirror of config::Encoder (/root/reference/src/config.rs:85-432)."""

    _fields_ = [
        ("block_size", C.c_int32),
        ("multithread", C.c_int32),
        ("workers", C.c_int32),
        ("use_leftside", C.c_int32),
        ("use_rightside", C.c_int32),
        ("use_midside", C.c_int32),
        ("use_constant", C.c_int32),
        ("use_fixed", C.c_int32),
        ("use_lpc", C.c_int32),
        ("fixed_max_order", C.c_int32),
        ("fixed_order_sel", C.c_int32),
        ("approx_ent_partitions", C.c_int32),
        ("lpc_order", C.c_int32),
        ("quant_precision", C.c_int32),
        ("use_direct_mse", C.c_int32),
        ("mae_optimization_steps", C.c_int32),
        ("window_type", C.c_int32),
        ("tukey_alpha", C.c_float),
        ("prc_max_parameter", C.c_int32),
        ("ext_lpc_order_search", C.c_int32),
        ("ext_lpc_precision_search", C.c_int32),
    ]


class SubframeInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("order", C.c_int32),
        ("bits_per_sample", C.c_int32),
        ("precision", C.c_int32),
        ("shift", C.c_int32),
        ("partition_order", C.c_int32),
        ("rice2", C.c_int32),
        ("reserved", C.c_int32),
        ("qlp", C.c_int16 * 32),
        ("rice_params", C.c_uint8 * MAX_RICE_PARTS),
        ("bits", C.c_uint64),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("channel_assignment", C.c_int32),
        ("block_size", C.c_int32),
        ("frame_number", C.c_uint32),
        ("frame_bytes", C.c_uint32),
        ("sub", SubframeInfo * MAX_CHANNELS),
    ]


class VariantTaps(C.Structure):
    _fields_ = [
        ("autocorr", C.c_double * (MAX_LPC_ORDER + 1)),
        ("lpc", C.c_double * MAX_LPC_ORDER),
        ("qlp", C.c_int16 * 32),
        ("qlp_order", C.c_int32),
        ("qlp_shift", C.c_int32),
        ("is_constant", C.c_int32),
        ("fixed_order", C.c_int32),
        ("fixed_est_bits", C.c_uint64 * 5),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("kernels_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("k_ingest_ms", C.c_float), ("k_analyze_ms", C.c_float), ("k_rice_ms", C.c_float),
        ("k_pack_ms", C.c_float), ("k_gather_ms", C.c_float),
        ("launches", C.c_uint64), ("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64),
        ("fused_frames", C.c_uint64), ("fallback_frames", C.c_uint64),
    ]


# every symbol include/flacenc_b200.h declares: name -> (restype, argtypes)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_szp = C.POINTER(C.c_size_t)
SYMBOLS = {
    "fb200_config_default": (None, [C.POINTER(Config)]),
    "fb200_config_verify": (C.c_int, [C.POINTER(Config)]),
    "fb200_device_count": (C.c_int, []),
    "fb200_create": (C.c_void_p, [C.POINTER(Config), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "fb200_destroy": (None, [C.c_void_p]),
    "fb200_max_frame_bytes": (C.c_size_t, [C.c_void_p]),
    "fb200_encode_interleaved": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_void_p, _szp, _szp]),
    "fb200_encode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p,
                                      C.c_size_t, C.c_void_p, _szp, _szp]),
    "fb200_encode_planar_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_uint32, C.c_void_p,
                                            C.c_size_t, _szp, C.POINTER(FrameInfo)]),
    "fb200_analyze": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.POINTER(VariantTaps), C.c_size_t, _szp]),
    "fb200_encode_stream": (C.c_int, [C.POINTER(Config), C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_size_t, _szp]),
    "fb200_encode_interleaved_sharded": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                                   C.c_void_p, C.c_size_t, C.c_void_p, _szp, _szp]),
    "fb200_encode_streams": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "fb200_md5": (None, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "fb200_pool_clear": (None, []),
    "fb200_debug_chunk_schedule": (C.c_size_t, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]),
    "fb200_debug_log2f": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_void_p]),
    "fb200_debug_irls_weight": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_void_p]),
    "fb200_debug_find_shift": (C.c_int, [C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "fb200_last_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "fb200_strerror": (C.c_char_p, [C.c_int]),
    "fb200_last_error": (C.c_char_p, [C.c_void_p]),
    "fb200_version": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first (__graft_entry__.build()); "
                "flacenc_rs_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
