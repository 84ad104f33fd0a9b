"""Host-side mirror of the reference's encode entry points, bound to the C ABI of libflacenc_b200.so.

  encode_with_fixed_block_size(config, src, block_size) -> Stream   /root/reference/src/coding.rs:645-695
  encode_fixed_size_frame(config, framebuf, frame_number, stream_info) -> Frame   src/coding.rs:581-606

Every call goes through ``_ffi.lib()`` (ctypes -> extern "C" -> CUDA kernels).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading
from collections import OrderedDict
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .config import Encoder as EncoderConfig, Verified
from .error import VerifyError, raise_for_code
from .source import FrameBuf, MemSource


@dataclass
class StreamInfo:
    """component::StreamInfo essentials (src/component/datatype.rs:435-560)."""
    sample_rate: int
    channels: int
    bits_per_sample: int

    @staticmethod
    def new(sample_rate: int, channels: int, bits_per_sample: int) -> "StreamInfo":
        if not 1 <= channels <= 8:
            raise VerifyError("channels", "must be in 1..=8")
        if sample_rate > 96000:
            raise VerifyError("sample_rate", "must be <= 96000")
        if not (8 <= bits_per_sample <= 25 and bits_per_sample % 4 in (0, 1)):
            raise VerifyError("bits_per_sample", "must be a multiple of 4 (or 4n + 1 for side-channel)")
        return StreamInfo(sample_rate, channels, bits_per_sample)


@dataclass
class SubFrame:
    """component::SubFrame as a decision record (what the Rust shim feeds to Constant::new /
    Verbatim::new / FixedLpc::new / Lpc::new + Residual::new)."""
    type: int
    order: int
    bits_per_sample: int
    precision: int
    shift: int
    partition_order: int
    rice2: bool
    qlp: List[int]
    rice_params: List[int]
    bits: int

    @staticmethod
    def from_info(s: _ffi.SubframeInfo) -> "SubFrame":
        coded = s.type in (_ffi.SF_FIXED, _ffi.SF_LPC)
        return SubFrame(s.type, s.order, s.bits_per_sample, s.precision, s.shift, s.partition_order, bool(s.rice2),
                        [s.qlp[i] for i in range(s.order)] if s.type == _ffi.SF_LPC else [],
                        [s.rice_params[i] for i in range(1 << s.partition_order)] if coded else [], s.bits)

    def count_bits(self) -> int:
        return self.bits


@dataclass
class Frame:
    """component::Frame with its precomputed bitstream (src/component/datatype.rs:820-828)."""
    channel_assignment: int
    block_size: int
    frame_number: int
    subframes: List[SubFrame]
    bitstream: bytes

    def count_bits(self) -> int:
        return len(self.bitstream) * 8


class Stream:
    """component::Stream: STREAMINFO + frames; ``write()`` yields the .flac bytes
    (src/component/bitrepr.rs:172-197)."""

    def __init__(self, data):
        self._data = data  # bytes, or the uint8 array the library filled (no copy until write() is asked for bytes)

    def write(self) -> bytes:
        if not isinstance(self._data, bytes):
            self._data = self._data.tobytes()
        return self._data

    def as_array(self) -> np.ndarray:
        """the stream bytes as a uint8 array, without a copy when the library's output buffer is still held"""
        return np.frombuffer(self._data, np.uint8) if isinstance(self._data, bytes) else self._data

    def count_bits(self) -> int:
        return len(self._data) * 8

    def __len__(self) -> int:
        return len(self._data)


class Context:
    """One stream format on one device (wraps fb200_create / fb200_destroy)."""

    def __init__(self, config: Verified, channels: int, bits_per_sample: int, sample_rate: int, block_size: int,
                 device: int = 0):
        self._lib = _ffi.lib()
        err = C.c_int(0)
        pod = config.pod
        self._h = self._lib.fb200_create(C.byref(pod), channels, bits_per_sample, sample_rate, block_size, device,
                                         C.byref(err))
        if not self._h:
            raise_for_code(err.value or _ffi.ERR_CUDA, "fb200_create failed (is a CUDA device visible?)")
        self.channels, self.bps, self.sample_rate, self.block_size = channels, bits_per_sample, sample_rate, block_size
        self.nvar = 4 if channels == 2 else channels

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _err(self) -> str:
        return (self._lib.fb200_last_error(self._h) or b"").decode()

    def max_frame_bytes(self) -> int:
        return self._lib.fb200_max_frame_bytes(self._h)

    def encode_interleaved(self, pcm, container_bytes: int, n_samples: int, first_frame_number: int = 0,
                           want_infos: bool = False, out: Optional[np.ndarray] = None):
        """fb200_encode_interleaved on a HOST buffer of packed little-endian interleaved PCM.
        Returns (frame bytes as a uint8 array view, frame sizes, infos or None)."""
        buf = np.ascontiguousarray(pcm).view(np.uint8).reshape(-1)
        n_frames = (n_samples + self.block_size - 1) // self.block_size
        cap = max(16, n_frames * self.max_frame_bytes())
        if out is None or len(out) < cap:
            out = np.empty(cap, np.uint8)
        sizes = np.zeros(max(n_frames, 1), np.uint32)
        infos = (_ffi.FrameInfo * max(n_frames, 1))() if want_infos else None
        nf, olen = C.c_size_t(0), C.c_size_t(0)
        rc = self._lib.fb200_encode_interleaved(self._h, buf.ctypes.data, container_bytes, n_samples,
                                                first_frame_number, out.ctypes.data, len(out), sizes.ctypes.data,
                                                C.cast(infos, C.c_void_p) if infos is not None else None,
                                                C.byref(nf), C.byref(olen))
        raise_for_code(rc, self._err())
        return out[: olen.value], sizes[: nf.value], infos

    def encode_device(self, d_pcm_ptr: int, container_bytes: int, n_samples: int, d_out_ptr: int, out_cap: int,
                      first_frame_number: int = 0, sizes: Optional[np.ndarray] = None):
        """fb200_encode_device: input and output already resident in device memory."""
        n_frames = (n_samples + self.block_size - 1) // self.block_size
        if sizes is None:
            sizes = np.zeros(max(n_frames, 1), np.uint32)
        nf, olen = C.c_size_t(0), C.c_size_t(0)
        rc = self._lib.fb200_encode_device(self._h, d_pcm_ptr, container_bytes, n_samples, first_frame_number,
                                           d_out_ptr, out_cap, sizes.ctypes.data, C.byref(nf), C.byref(olen))
        raise_for_code(rc, self._err())
        return olen.value, sizes[: nf.value]

    def encode_planar_frame(self, planar: np.ndarray, n: int, frame_number: int):
        pl = np.ascontiguousarray(planar, np.int32)
        stride = pl.shape[1] if pl.ndim == 2 else len(pl) // self.channels
        out = np.empty(self.max_frame_bytes(), np.uint8)
        olen = C.c_size_t(0)
        info = _ffi.FrameInfo()
        rc = self._lib.fb200_encode_planar_frame(self._h, pl.ctypes.data_as(C.POINTER(C.c_int32)), stride, n,
                                                 frame_number, out.ctypes.data, len(out), C.byref(olen), C.byref(info))
        raise_for_code(rc, self._err())
        return out[: olen.value].tobytes(), info

    def analyze(self, pcm, container_bytes: int, n_samples: int):
        buf = np.ascontiguousarray(pcm).view(np.uint8).reshape(-1)
        n_frames = (n_samples + self.block_size - 1) // self.block_size
        taps = (_ffi.VariantTaps * max(n_frames * self.nvar, 1))()
        nv = C.c_size_t(0)
        rc = self._lib.fb200_analyze(self._h, buf.ctypes.data, container_bytes, n_samples, taps, n_frames * self.nvar,
                                     C.byref(nv))
        raise_for_code(rc, self._err())
        return taps, nv.value

    def timing(self) -> _ffi.Timing:
        t = _ffi.Timing()
        self._lib.fb200_last_timing(self._h, C.byref(t))
        return t


def pack_samples(samples: np.ndarray, bytes_per_sample: int) -> np.ndarray:
    """int32 samples -> packed little-endian bytes of ``bytes_per_sample`` each (what a WAV reader hands to
    Fill::fill_le_bytes, /root/reference/flacenc-bin/src/source.rs:86-133)."""
    x = np.ascontiguousarray(samples, np.int32).reshape(-1)
    if bytes_per_sample == 4:
        return x.astype("<i4").view(np.uint8)
    if bytes_per_sample == 2:
        return x.astype("<i2").view(np.uint8)
    if bytes_per_sample == 3:
        return np.ascontiguousarray(x.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)
    if bytes_per_sample == 1:
        return x.astype(np.int8).view(np.uint8)
    raise ValueError(bytes_per_sample)


def encode_with_fixed_block_size(config: Verified, src: MemSource, block_size: int,
                                 devices: Optional[Sequence[int]] = None) -> Stream:
    """``flacenc::encode_with_fixed_block_size`` (src/coding.rs:645-695): Source -> Stream.
    Frames are sharded by contiguous range over ``devices`` (default: device 0), MD5 and STREAMINFO
    are assembled on the host (fb200_encode_stream)."""
    if not isinstance(config, Verified):
        raise TypeError("config must be Verified (use Encoder().into_verified())")
    lib = _ffi.lib()
    ch, bps, rate = src.channels(), src.bits_per_sample(), src.sample_rate()
    n = len(src)
    # the int32 samples go to the library as they are (container 4 = Fill::fill_interleaved); its MD5 thread packs them
    # to ceil(bps / 8) bytes piece by piece (src/source.rs:406-418), so no packed copy of the stream is ever made here
    cb = 4
    # (nothing is truncated on the way: FrameBuf::verify_samples, src/source.rs:262-275, is the ingest kernel's range
    # check on the true int32 values -> VerifyError)
    samples = np.ascontiguousarray(src.as_interleaved(), np.int32)
    pcm = samples.reshape(-1)
    n_frames = (n + block_size - 1) // block_size
    cap = 64 + n_frames * (32 + ch * ((block_size * (bps + 1) + 7) // 8 + 2))
    out = np.empty(cap, np.uint8)
    olen = C.c_size_t(0)
    devs = list(devices) if devices else [0]
    dev_arr = (C.c_int * len(devs))(*devs)
    pod = config.pod
    rc = lib.fb200_encode_stream(C.byref(pod), pcm.ctypes.data, cb, n, ch, bps, rate, block_size, dev_arr, len(devs),
                                 out.ctypes.data, cap, C.byref(olen))
    raise_for_code(rc, "fb200_encode_stream: " + (lib.fb200_last_error(None) or b"").decode())
    return Stream(out[: olen.value])


def encode_streams_with_fixed_block_size(config: Verified, sources: Sequence[MemSource], block_size: int,
                                         devices: Optional[Sequence[int]] = None) -> List[Stream]:
    """A batch of sources of ONE format (fb200_encode_streams): what calling ``encode_with_fixed_block_size`` per file
    does, with the MD5 of every stream on its own host thread (src/par.rs:196-277) while the devices are fed stream by
    stream."""
    if not isinstance(config, Verified):
        raise TypeError("config must be Verified (use Encoder().into_verified())")
    if not sources:
        return []
    lib = _ffi.lib()
    ch, bps, rate = sources[0].channels(), sources[0].bits_per_sample(), sources[0].sample_rate()
    pcms, outs, ns = [], [], []
    for src in sources:
        if (src.channels(), src.bits_per_sample(), src.sample_rate()) != (ch, bps, rate):
            raise ValueError("all sources of a batch must share channels, bits_per_sample and sample_rate")
        x = np.ascontiguousarray(src.as_interleaved(), np.int32).reshape(-1)  # range-checked on the device
        n = len(src)
        n_frames = (n + block_size - 1) // block_size
        pcms.append(x)
        ns.append(n)
        outs.append(np.empty(64 + n_frames * (32 + ch * ((block_size * (bps + 1) + 7) // 8 + 2)), np.uint8))
    k = len(sources)
    pcm_arr = (C.c_void_p * k)(*[x.ctypes.data for x in pcms])
    out_arr = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
    n_arr = (C.c_uint64 * k)(*ns)
    cap_arr = (C.c_size_t * k)(*[len(o) for o in outs])
    len_arr = (C.c_size_t * k)()
    rc_arr = (C.c_int * k)()
    devs = list(devices) if devices else [0]
    dev_arr = (C.c_int * len(devs))(*devs)
    pod = config.pod
    rc = lib.fb200_encode_streams(C.byref(pod), k, pcm_arr, n_arr, 4, ch, bps, rate, block_size, dev_arr, len(devs), out_arr,
                                  cap_arr, len_arr, rc_arr)
    raise_for_code(rc, "fb200_encode_streams: " + (lib.fb200_last_error(None) or b"").decode())
    return [Stream(outs[i][: len_arr[i]]) for i in range(k)]


def encode_interleaved_sharded(contexts: Sequence[Context], pcm, container_bytes: int, n_samples: int,
                               first_frame_number: int = 0, out: Optional[np.ndarray] = None):
    """fb200_encode_interleaved_sharded: one batch over the devices of ``contexts`` by frame range (src/par.rs:355-449);
    returns (frame bytes view, frame sizes).  ``contexts[0].timing()`` describes the call."""
    c0 = contexts[0]
    buf = np.ascontiguousarray(pcm).view(np.uint8).reshape(-1)
    n_frames = (n_samples + c0.block_size - 1) // c0.block_size
    cap = max(16, n_frames * c0.max_frame_bytes())
    if out is None or len(out) < cap:
        out = np.empty(cap, np.uint8)
    sizes = np.zeros(max(n_frames, 1), np.uint32)
    handles = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    nf, olen = C.c_size_t(0), C.c_size_t(0)
    rc = c0._lib.fb200_encode_interleaved_sharded(handles, len(contexts), buf.ctypes.data, container_bytes, n_samples,
                                                  first_frame_number, out.ctypes.data, len(out), sizes.ctypes.data,
                                                  C.byref(nf), C.byref(olen))
    raise_for_code(rc, c0._err())
    return out[: olen.value], sizes[: nf.value]


# The reference calls encode_fixed_size_frame in its inner loop (src/par.rs:384-389) and keeps all scratch in
# thread-local `reusable!` storage (src/lib.rs:92-116).  The mirror keeps one device context per thread and
# (config, format, block size, device): device buffers, window and CRC tables are built once, not per frame.
_CTX_CACHE = threading.local()
_CTX_CACHE_MAX = 8


def _cached_context(config: Verified, channels: int, bps: int, rate: int, block_size: int, device: int) -> Context:
    cache = getattr(_CTX_CACHE, "ctxs", None)
    if cache is None:
        cache = _CTX_CACHE.ctxs = OrderedDict()
    key = (bytes(config.pod), channels, bps, rate, block_size, device)
    ctx = cache.get(key)
    if ctx is None:
        ctx = cache[key] = Context(config, channels, bps, rate, block_size, device)
        while len(cache) > _CTX_CACHE_MAX:
            cache.popitem(last=False)[1].close()
    else:
        cache.move_to_end(key)
    return ctx


def encode_fixed_size_frame(config: Verified, framebuf: FrameBuf, frame_number: int, stream_info: StreamInfo,
                            device: int = 0) -> Frame:
    """``flacenc::encode_fixed_size_frame`` (src/coding.rs:581-606): FrameBuf -> Frame."""
    if not isinstance(config, Verified):
        raise TypeError("config must be Verified (use Encoder().into_verified())")
    if not 0 <= frame_number < (1 << 31):
        raise VerifyError("encode_fixed_size_frame (frame_number)", "must be < 2^31")
    ctx = _cached_context(config, stream_info.channels, stream_info.bits_per_sample, stream_info.sample_rate,
                          framebuf.size(), device)
    planar = framebuf.samples.reshape(framebuf.channels(), framebuf.size())
    data, info = ctx.encode_planar_frame(planar, framebuf.filled_size(), frame_number)
    subs = [SubFrame.from_info(info.sub[c]) for c in range(stream_info.channels)]
    return Frame(info.channel_assignment, info.block_size, info.frame_number, subs, data)
