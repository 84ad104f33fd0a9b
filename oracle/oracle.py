"""ctypes binding of the CPU oracle (oracle/libflacenc_oracle.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package (flacenc_rs_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libflacenc_oracle.so")

MAX_LPC_ORDER = 24
MAX_RICE_PARTS = 256
MAX_CHANNELS = 8


class Config(C.Structure):
    """POD mirror of config::Encoder; same layout as fb200_config (include/flacenc_b200.h)."""

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


class SubFrame(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("order", C.c_int32),
        ("bps", C.c_int32),
        ("n", C.c_int32),
        ("precision", C.c_int32),
        ("shift", C.c_int32),
        ("part_order", C.c_int32),
        ("rice2", C.c_int32),
        ("qlp", C.c_int16 * 32),
        ("rice_params", C.c_uint8 * MAX_RICE_PARTS),
        ("sum_quotients", C.c_uint64),
        ("code_bits", C.c_uint64),
        ("bits", C.c_uint64),
        ("residual", C.POINTER(C.c_int32)),
        ("samples", C.POINTER(C.c_int32)),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("channels", C.c_int32),
        ("n", C.c_int32),
        ("bps", C.c_int32),
        ("sample_rate", C.c_int32),
        ("ch_assignment", C.c_int32),
        ("frame_number", C.c_uint32),
        ("sub", SubFrame * MAX_CHANNELS),
        ("ms_buf", C.POINTER(C.c_int32)),
    ]


class StreamInfo(C.Structure):
    _fields_ = [
        ("min_block", C.c_uint32),
        ("max_block", C.c_uint32),
        ("min_frame", C.c_uint32),
        ("max_frame", C.c_uint32),
        ("sample_rate", C.c_uint32),
        ("channels", C.c_uint32),
        ("bps", C.c_uint32),
        ("total_samples", C.c_uint64),
        ("md5", C.c_uint8 * 16),
        ("n_frames", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("flacenc_oracle.c", "flacenc_decoder.c", "flacenc_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libflacenc_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32p = C.POINTER(C.c_int32)
        u32p = C.POINTER(C.c_uint32)
        u8p = C.POINTER(C.c_uint8)
        f32p = C.POINTER(C.c_float)
        f64p = C.POINTER(C.c_double)
        i16p = C.POINTER(C.c_int16)
        cfgp = C.POINTER(Config)
        sig = {
            "fo_config_default": (None, [cfgp]),
            "fo_config_verify": (C.c_int, [cfgp]),
            "fo_window_weights": (None, [C.c_int, C.c_float, C.c_int, f32p]),
            "fo_fill_windowed_signal": (None, [i32p, f32p, C.c_int, f32p]),
            "fo_auto_correlation_f64": (None, [C.c_int, f32p, C.c_int, f64p]),
            "fo_auto_correlation_f32": (None, [C.c_int, f32p, C.c_int, f32p]),
            "fo_levinson_f64": (None, [f64p, f64p, C.c_int, f64p]),
            "fo_levinson_f32": (None, [f32p, f32p, C.c_int, f32p]),
            "fo_find_shift": (C.c_int, [f64p, C.c_int, C.c_int]),
            "fo_log2f_bits": (None, [C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]),
            "fo_find_shift_each": (None, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
            "fo_quantize_parameters": (C.c_int, [f64p, C.c_int, C.c_int, i16p, C.POINTER(C.c_int)]),
            "fo_compute_error": (None, [i16p, C.c_int, C.c_int, i32p, C.c_int, i32p]),
            "fo_lpc_from_autocorr": (None, [i32p, C.c_int, C.c_int, C.c_float, C.c_int, f64p, f64p]),
            "fo_lagged_outer_prod_sum": (None, [C.c_int, f32p, C.c_int, f64p]),
            "fo_solve_sym": (C.c_int, [f64p, C.c_int, f64p]),
            "fo_lpc_with_direct_mse": (None, [i32p, C.c_int, C.c_int, C.c_float, C.c_int, f64p, f64p, f64p]),
            "fo_compute_raw_errors": (None, [i32p, C.c_int, f64p, C.c_int, f32p]),
            "fo_ext_lpc_orders": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
            "fo_irls_weight": (C.c_float, [C.c_float, C.c_float]),
            "fo_irls_weight_bits": (None, [C.c_uint32, C.c_uint64, C.c_float, C.c_int, C.c_void_p]),
            "fo_lpc_with_irls_mae": (None, [i32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, f64p, f32p]),
            "fo_encode_signbit": (C.c_uint32, [C.c_int32]),
            "fo_decode_signbit": (C.c_int32, [C.c_uint32]),
            "fo_finest_partition_order": (C.c_int, [C.c_int, C.c_int]),
            "fo_bit_table_from_errors": (None, [u32p, C.c_int, C.c_uint32, u32p]),
            "fo_bit_table_minimizer": (None, [u32p, C.c_int, C.POINTER(C.c_int), u32p]),
            "fo_bit_table_merge": (None, [u32p, u32p, C.c_uint32, u32p]),
            "fo_find_partitioned_rice_parameter": (C.c_int, [i32p, C.c_int, C.c_int, C.c_int, u8p, C.POINTER(C.c_uint64)]),
            "fo_fixed_lpc_errors": (None, [i32p, C.c_int, i32p]),
            "fo_estimate_entropy": (C.c_uint64, [i32p, C.c_int, C.c_int, C.c_int]),
            "fo_select_order": (C.c_int, [C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]),
            "fo_encode_subframe": (None, [cfgp, i32p, C.c_int, C.c_int, C.POINTER(SubFrame)]),
            "fo_encode_frame": (C.c_int, [cfgp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(Frame)]),
            "fo_frame_free": (None, [C.POINTER(Frame)]),
            "fo_crc8": (C.c_uint8, [u8p, C.c_size_t]),
            "fo_crc16": (C.c_uint16, [u8p, C.c_size_t]),
            "fo_encode_utf8like": (C.c_int, [C.c_uint64, u8p]),
            "fo_frame_header_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, u8p]),
            "fo_frame_count_bits": (C.c_uint64, [C.POINTER(Frame)]),
            "fo_subframe_write": (C.c_int64, [C.POINTER(SubFrame), u8p, C.c_size_t]),
            "fo_frame_write": (C.c_int64, [C.POINTER(Frame), u8p, C.c_size_t]),
            "fo_encode_frames": (C.c_int64, [cfgp, i32p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, u8p, C.c_size_t, u32p]),
            "fo_encode_stream": (C.c_int64, [cfgp, i32p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_size_t]),
            "fo_md5": (None, [u8p, C.c_size_t, u8p]),
            "fo_md5_of_samples": (None, [i32p, C.c_size_t, C.c_int, u8p]),
            "fo_decode_stream": (C.c_int64, [u8p, C.c_size_t, C.POINTER(StreamInfo), i32p, C.c_uint64]),
            "fo_decode_frames": (C.c_int64, [u8p, C.c_size_t, C.c_int, C.c_int, i32p, C.c_uint64, C.POINTER(C.c_uint64)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


# ------------------------------------------------------------------ helpers


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def default_config(**overrides) -> Config:
    cfg = Config()
    lib().fo_config_default(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def window_weights(window_type: int, alpha: float, n: int) -> np.ndarray:
    out = np.zeros(n, np.float32)
    lib().fo_window_weights(window_type, alpha, n, _p(out, C.c_float))
    return out


def auto_correlation(order: int, signal: np.ndarray, dtype=np.float64) -> np.ndarray:
    sig = np.ascontiguousarray(signal, np.float32)
    dest = np.zeros(order, dtype)
    if dtype == np.float64:
        lib().fo_auto_correlation_f64(order, _p(sig, C.c_float), len(sig), _p(dest, C.c_double))
    else:
        lib().fo_auto_correlation_f32(order, _p(sig, C.c_float), len(sig), _p(dest, C.c_float))
    return dest


def levinson(coefs, ys, dtype=np.float64) -> np.ndarray:
    c = np.ascontiguousarray(coefs, dtype)
    y = np.ascontiguousarray(ys, dtype)
    out = np.zeros(len(y), dtype)
    if dtype == np.float64:
        lib().fo_levinson_f64(_p(c, C.c_double), _p(y, C.c_double), len(y), _p(out, C.c_double))
    else:
        lib().fo_levinson_f32(_p(c, C.c_float), _p(y, C.c_float), len(y), _p(out, C.c_float))
    return out


def log2f_bits(first: int, count: int, threads: int = 8) -> np.ndarray:
    """bits of the host libm's log2f over the float bit patterns [first, first + count)"""
    out = np.empty(count, np.uint32)
    lib().fo_log2f_bits(first, count, threads, out.ctypes.data)
    return out


def find_shift_each(values, precision: int) -> np.ndarray:
    v = np.ascontiguousarray(values, np.float64)
    out = np.empty(len(v), np.int32)
    lib().fo_find_shift_each(v.ctypes.data, len(v), precision, out.ctypes.data)
    return out


def find_shift(coefs, precision: int) -> int:
    c = np.ascontiguousarray(coefs, np.float64)
    return lib().fo_find_shift(_p(c, C.c_double), len(c), precision)


def quantize_parameters(coefs, precision: int):
    c = np.ascontiguousarray(coefs, np.float64)
    q = np.zeros(MAX_LPC_ORDER, np.int16)
    shift = C.c_int(0)
    order = lib().fo_quantize_parameters(_p(c, C.c_double), len(c), precision, _p(q, C.c_int16), C.byref(shift))
    return q[:order].copy(), order, shift.value


def compute_error(q, shift: int, signal) -> np.ndarray:
    qq = np.zeros(MAX_LPC_ORDER, np.int16)
    qq[: len(q)] = q
    s = np.ascontiguousarray(signal, np.int32)
    e = np.zeros(len(s), np.int32)
    lib().fo_compute_error(_p(qq, C.c_int16), len(q), shift, _p(s, C.c_int32), len(s), _p(e, C.c_int32))
    return e


def lpc_from_autocorr(signal, window_type: int, alpha: float, lpc_order: int):
    s = np.ascontiguousarray(signal, np.int32)
    coefs = np.zeros(max(lpc_order, 1), np.float64)
    corr = np.zeros(lpc_order + 1, np.float64)
    lib().fo_lpc_from_autocorr(_p(s, C.c_int32), len(s), window_type, alpha, lpc_order, _p(coefs, C.c_double), _p(corr, C.c_double))
    return coefs[:lpc_order].copy(), corr


def lagged_outer_prod_sum(order: int, signal) -> np.ndarray:
    x = np.ascontiguousarray(signal, np.float32)
    dest = np.zeros((order, order), np.float64)
    lib().fo_lagged_outer_prod_sum(order, _p(x, C.c_float), len(x), _p(dest, C.c_double))
    return dest


def solve_sym(mat, v):
    """LpcFloat::solve_sym_mut: (ok, solution)"""
    m = np.ascontiguousarray(mat, np.float64)
    x = np.ascontiguousarray(v, np.float64).copy()
    ok = lib().fo_solve_sym(_p(m, C.c_double), len(x), _p(x, C.c_double))
    return bool(ok), x


def lpc_with_direct_mse(signal, window_type: int, alpha: float, lpc_order: int):
    """(coefs, autocorrelation lags 0..order, covariance matrix) of the `experimental` covariance-method estimator"""
    s = np.ascontiguousarray(signal, np.int32)
    coefs = np.zeros(max(lpc_order, 1), np.float64)
    corr = np.zeros(lpc_order + 1, np.float64)
    covar = np.zeros((max(lpc_order, 1), max(lpc_order, 1)), np.float64)
    lib().fo_lpc_with_direct_mse(_p(s, C.c_int32), len(s), window_type, alpha, lpc_order, _p(coefs, C.c_double),
                                 _p(corr, C.c_double), _p(covar, C.c_double))
    return coefs[:lpc_order].copy(), corr, covar[:lpc_order, :lpc_order] if False else covar.reshape(-1)[: lpc_order * lpc_order].reshape(lpc_order, lpc_order)


def compute_raw_errors(signal, coefs) -> np.ndarray:
    """src/lpc.rs:606-618: f32 "prediction - signal" of unquantised coefficients (zeros before the order)"""
    s = np.ascontiguousarray(signal, np.int32)
    c = np.ascontiguousarray(coefs, np.float64)
    e = np.zeros(len(s), np.float32)
    lib().fo_compute_raw_errors(_p(s, C.c_int32), len(s), _p(c, C.c_double), len(c), _p(e, C.c_float))
    return e


def irls_weight(err: float, normalizer: float) -> float:
    return float(lib().fo_irls_weight(err, normalizer))


def ext_lpc_orders(lpc_order: int, k: int):
    """the LPC orders the ext_lpc_order_search extension tries (lpc_order first)"""
    buf = (C.c_int * 16)()
    n = lib().fo_ext_lpc_orders(lpc_order, k, buf)
    return list(buf[:n])


def irls_weight_bits(first: int, count: int, normalizer: float, threads: int = 8) -> np.ndarray:
    """bits of the IRLS weight (host libm powf) of the raw errors with float bit patterns [first, first + count)"""
    out = np.empty(count, np.uint32)
    lib().fo_irls_weight_bits(first, count, normalizer, threads, out.ctypes.data)
    return out


def lpc_with_irls_mae(signal, window_type: int, alpha: float, lpc_order: int, steps: int):
    """(coefs, the steps + 1 sums of |raw error|) of the `experimental` IRLS-MAE estimator (src/lpc.rs:814-850)"""
    s = np.ascontiguousarray(signal, np.int32)
    coefs = np.zeros(max(lpc_order, 1), np.float64)
    sums = np.zeros(steps + 1, np.float32)
    lib().fo_lpc_with_irls_mae(_p(s, C.c_int32), len(s), window_type, alpha, lpc_order, steps, _p(coefs, C.c_double),
                               _p(sums, C.c_float))
    return coefs[:lpc_order].copy(), sums


def bit_table_from_errors(errors, offset: int) -> np.ndarray:
    e = np.ascontiguousarray(errors, np.uint32)
    t = np.zeros(32, np.uint32)
    lib().fo_bit_table_from_errors(_p(e, C.c_uint32), len(e), offset, _p(t, C.c_uint32))
    return t


def bit_table_minimizer(table, max_p: int):
    t = np.ascontiguousarray(table, np.uint32)
    p = C.c_int(0)
    bits = C.c_uint32(0)
    lib().fo_bit_table_minimizer(_p(t, C.c_uint32), max_p, C.byref(p), C.byref(bits))
    return p.value, bits.value


def bit_table_merge(a, b, offset: int) -> np.ndarray:
    aa = np.ascontiguousarray(a, np.uint32)
    bb = np.ascontiguousarray(b, np.uint32)
    out = np.zeros(32, np.uint32)
    lib().fo_bit_table_merge(_p(aa, C.c_uint32), _p(bb, C.c_uint32), offset, _p(out, C.c_uint32))
    return out


def find_partitioned_rice_parameter(signal, warmup: int, max_p: int):
    s = np.ascontiguousarray(signal, np.int32)
    ps = np.zeros(MAX_RICE_PARTS, np.uint8)
    code_bits = C.c_uint64(0)
    order = lib().fo_find_partitioned_rice_parameter(_p(s, C.c_int32), len(s), warmup, max_p, _p(ps, C.c_uint8), C.byref(code_bits))
    return order, ps[: 1 << order].copy(), code_bits.value


def fixed_lpc_errors(signal) -> np.ndarray:
    s = np.ascontiguousarray(signal, np.int32)
    out = np.zeros((5, len(s)), np.int32)
    lib().fo_fixed_lpc_errors(_p(s, C.c_int32), len(s), _p(out, C.c_int32))
    return out


def estimate_entropy(errors, warmup: int, partitions: int) -> int:
    e = np.ascontiguousarray(errors, np.int32)
    return lib().fo_estimate_entropy(_p(e, C.c_int32), len(e), warmup, partitions)


def select_order(order_sel: int, partitions: int, max_p: int, errors2d, bps: int, baseline_bits: int):
    e = np.ascontiguousarray(errors2d, np.int32)
    bits = C.c_uint64(0)
    order = lib().fo_select_order(order_sel, partitions, max_p, _p(e, C.c_int32), e.shape[0], e.shape[1], bps, baseline_bits, C.byref(bits))
    return order, bits.value


def subframe_record(sf: SubFrame) -> dict:
    d = {
        "type": sf.type, "order": sf.order, "bps": sf.bps, "n": sf.n, "bits": sf.bits,
    }
    if sf.type == 3:
        d.update(precision=sf.precision, shift=sf.shift, qlp=[sf.qlp[i] for i in range(sf.order)])
    if sf.type in (2, 3):
        d.update(part_order=sf.part_order, rice2=sf.rice2, code_bits=sf.code_bits,
                 rice_params=[sf.rice_params[i] for i in range(1 << sf.part_order)],
                 residual=np.ctypeslib.as_array(sf.residual, shape=(sf.n,)).copy())
    return d


def encode_subframe(cfg: Config, samples, bps: int) -> dict:
    s = np.ascontiguousarray(samples, np.int32)
    fr = Frame()  # reuse Frame as an owner so fo_frame_free releases the residual
    lib().fo_encode_subframe(C.byref(cfg), _p(s, C.c_int32), len(s), bps, C.byref(fr.sub[0]))
    rec = subframe_record(fr.sub[0])
    lib().fo_frame_free(C.byref(fr))
    return rec


def encode_frame(cfg: Config, planar, bps: int, sample_rate: int, frame_number: int, n: int | None = None):
    """encode_fixed_size_frame on a planar (channels, stride) int32 array. Returns (bytes, record)."""
    pl = np.ascontiguousarray(planar, np.int32)
    channels, stride = pl.shape
    if n is None:
        n = stride
    fr = Frame()
    rc = lib().fo_encode_frame(C.byref(cfg), _p(pl, C.c_int32), channels, stride, n, bps, sample_rate, frame_number, C.byref(fr))
    if rc:
        lib().fo_frame_free(C.byref(fr))
        raise ValueError("VerifyError (oracle)")
    bits = lib().fo_frame_count_bits(C.byref(fr))
    buf = np.zeros(bits // 8, np.uint8)
    nbytes = lib().fo_frame_write(C.byref(fr), _p(buf, C.c_uint8), len(buf))
    assert nbytes == len(buf), (nbytes, len(buf))
    rec = {
        "ch_assignment": fr.ch_assignment,
        "bits": bits,
        "subframes": [subframe_record(fr.sub[ch]) for ch in range(channels)],
    }
    lib().fo_frame_free(C.byref(fr))
    return buf.tobytes(), rec


def encode_frames(cfg: Config, interleaved, channels: int, bps: int, sample_rate: int, block_size: int,
                  first_frame_number: int = 0, nthreads: int = 1):
    """Frames of a stream (no stream header). interleaved: int32 array of n*channels samples."""
    x = np.ascontiguousarray(interleaved, np.int32).reshape(-1)
    n = len(x) // channels
    n_frames = (n + block_size - 1) // block_size
    cap = 64 + n_frames * (32 + channels * (block_size * 4 + 64))
    out = np.zeros(cap, np.uint8)
    sizes = np.zeros(max(n_frames, 1), np.uint32)
    total = lib().fo_encode_frames(C.byref(cfg), _p(x, C.c_int32), n, channels, bps, sample_rate, block_size,
                                   first_frame_number, nthreads, _p(out, C.c_uint8), cap, _p(sizes, C.c_uint32))
    if total == -1:
        raise ValueError("VerifyError (oracle)")
    if total < 0:
        raise RuntimeError(f"oracle encode_frames failed: {total}")
    return out[:total].tobytes(), sizes[:n_frames].copy()


def encode_stream(cfg: Config, interleaved, channels: int, bps: int, sample_rate: int, block_size: int,
                  nthreads: int = 1) -> bytes:
    x = np.ascontiguousarray(interleaved, np.int32).reshape(-1)
    n = len(x) // channels
    n_frames = (n + block_size - 1) // block_size
    cap = 128 + n_frames * (32 + channels * (block_size * 4 + 64))
    out = np.zeros(cap, np.uint8)
    total = lib().fo_encode_stream(C.byref(cfg), _p(x, C.c_int32), n, channels, bps, sample_rate, block_size,
                                   nthreads, _p(out, C.c_uint8), cap)
    if total == -1:
        raise ValueError("VerifyError (oracle)")
    if total < 0:
        raise RuntimeError(f"oracle encode_stream failed: {total}")
    return out[:total].tobytes()


def md5(data: bytes) -> bytes:
    d = np.frombuffer(data, np.uint8)
    out = np.zeros(16, np.uint8)
    lib().fo_md5(_p(d, C.c_uint8) if len(d) else None, len(d), _p(out, C.c_uint8))
    return out.tobytes()


def md5_of_samples(interleaved, bytes_per_sample: int) -> bytes:
    x = np.ascontiguousarray(interleaved, np.int32).reshape(-1)
    out = np.zeros(16, np.uint8)
    lib().fo_md5_of_samples(_p(x, C.c_int32), len(x), bytes_per_sample, _p(out, C.c_uint8))
    return out.tobytes()


def decode_stream(data: bytes):
    """Independent decoder. Returns (interleaved int32 array (n, channels), StreamInfo)."""
    d = np.frombuffer(data, np.uint8)
    info = StreamInfo()
    n = lib().fo_decode_stream(_p(d, C.c_uint8), len(d), C.byref(info), None, 0)
    if n < 0:
        raise ValueError(f"decode error {n}")
    out = np.zeros((max(n, 1), info.channels), np.int32)
    n2 = lib().fo_decode_stream(_p(d, C.c_uint8), len(d), C.byref(info), _p(out, C.c_int32), n)
    assert n2 == n
    return out[:n], info


def decode_frames(data: bytes, channels: int, bps: int):
    d = np.frombuffer(data, np.uint8)
    nf = C.c_uint64(0)
    n = lib().fo_decode_frames(_p(d, C.c_uint8), len(d), channels, bps, None, 0, C.byref(nf))
    if n < 0:
        raise ValueError(f"decode error {n}")
    out = np.zeros((max(n, 1), channels), np.int32)
    lib().fo_decode_frames(_p(d, C.c_uint8), len(d), channels, bps, _p(out, C.c_int32), n, C.byref(nf))
    return out[:n], nf.value
