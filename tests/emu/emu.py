"""ctypes binding of the CPU emulation of the CUDA kernel bodies (tests/emu/libfb_emu.so).

TEST INFRASTRUCTURE: used by `-m "not gpu"` tests to check kernel logic against the oracle without a
GPU.  Shares the ctypes structures of the real C ABI (flacenc_rs_b200._ffi)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from flacenc_rs_b200 import _ffi as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfb_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE, "libfb_emu.so"], check=True, capture_output=True)
        L = C.CDLL(_LIB)
        L.fbemu_log2f.restype = C.c_float
        L.fbemu_log2f.argtypes = [C.c_float]
        L.fbemu_log2f_sweep.restype = C.c_ulonglong
        L.fbemu_log2f_sweep.argtypes = [C.c_uint32, C.c_ulonglong, C.c_int, C.POINTER(C.c_uint32)]
        L.fbemu_irls_weight_bits.restype = None
        L.fbemu_irls_weight_bits.argtypes = [C.c_uint32, C.c_ulonglong, C.c_float, C.c_int, C.c_void_p]
        L.fbemu_find_shift.restype = C.c_int
        L.fbemu_find_shift.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
        L.fbemu_config_default.argtypes = [C.POINTER(F.Config)]
        L.fbemu_config_verify.argtypes = [C.POINTER(F.Config)]
        L.fbemu_frame_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_uint8)]
        L.fbemu_encode_interleaved.argtypes = [
            C.POINTER(F.Config), C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
            C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(F.FrameInfo), C.POINTER(C.c_size_t),
            C.POINTER(C.c_size_t), C.c_int]
        L.fbemu_encode_planar_frame.argtypes = [
            C.POINTER(F.Config), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
            C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(F.FrameInfo)]
        L.fbemu_analyze.argtypes = [
            C.POINTER(F.Config), C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(F.VariantTaps), C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


def default_config(**kw) -> F.Config:
    cfg = F.Config()
    lib().fbemu_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def encode_interleaved(cfg, pcm_bytes: np.ndarray, container_bytes: int, n_samples: int, channels: int, bps: int,
                       rate: int, block_size: int, first_frame: int = 0, want_infos: bool = False):
    """pcm_bytes: uint8 array of packed little-endian interleaved samples. Returns (rc, bytes, sizes, infos)."""
    pcm = np.ascontiguousarray(pcm_bytes).view(np.uint8)
    n_frames = (n_samples + block_size - 1) // block_size
    cap = 64 + n_frames * (64 + channels * (block_size * 4 + 16))
    out = np.zeros(cap, np.uint8)
    sizes = np.zeros(max(n_frames, 1), np.uint32)
    infos = (F.FrameInfo * max(n_frames, 1))() if want_infos else None
    nf = C.c_size_t(0)
    olen = C.c_size_t(0)
    rc = lib().fbemu_encode_interleaved(
        C.byref(cfg), pcm.ctypes.data_as(C.c_void_p), container_bytes, n_samples, channels, bps, rate, block_size,
        first_frame, out.ctypes.data_as(C.POINTER(C.c_uint8)), cap, sizes.ctypes.data_as(C.POINTER(C.c_uint32)),
        infos, C.byref(nf), C.byref(olen), 0)
    return rc, out[: olen.value].tobytes() if rc == 0 else b"", sizes[: nf.value].copy(), infos


def analyze(cfg, pcm_bytes: np.ndarray, container_bytes: int, n_samples: int, channels: int, bps: int, rate: int,
            block_size: int):
    pcm = np.ascontiguousarray(pcm_bytes).view(np.uint8)
    n_frames = (n_samples + block_size - 1) // block_size
    nvar = 4 if channels == 2 else channels
    taps = (F.VariantTaps * max(n_frames * nvar, 1))()
    nv = C.c_size_t(0)
    rc = lib().fbemu_analyze(C.byref(cfg), pcm.ctypes.data_as(C.c_void_p), container_bytes, n_samples, channels, bps,
                             rate, block_size, taps, n_frames * nvar, C.byref(nv))
    return rc, taps, nv.value


def set_force_generic(flag: bool) -> None:
    """True: run only the generic K2/K3 kernels; False (default): the fused kernel when eligible, like the library."""
    lib().fbemu_set_force_generic(1 if flag else 0)


def launch_geometry(channels: int, bps: int, rate: int, block_size: int, n_samples: int):
    """(frames, lane slots of the analysis launch, variants before an isolated last frame, odd mode, staged rows)"""
    out = (C.c_uint32 * 4)()
    L = lib()
    L.fbemu_launch_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_uint32)]
    L.fbemu_launch_geometry.restype = C.c_int
    frames = L.fbemu_launch_geometry(channels, bps, rate, block_size, n_samples, out)
    return frames, out[0], out[1], out[2], out[3]


def set_kp_pairs(flag: bool) -> None:
    """False: the pack kernel always stages planes from the planar store (default True: PCM pairs when the format allows)."""
    lib().fbemu_set_kp_pairs(1 if flag else 0)


def fused_counts(reset: bool = True):
    """(frames encoded by the fused kernel, frames it handed to the generic kernels)"""
    out = (C.c_ulonglong * 2)()
    lib().fbemu_fused_counts(out, 1 if reset else 0)
    return list(out)


def mode_counts(reset: bool = True):
    out = (C.c_ulonglong * 4)()
    lib().fbemu_mode_counts(out, 1 if reset else 0)
    return list(out)
