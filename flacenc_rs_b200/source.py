"""Host-side mirror of the reference's ``source`` module (/root/reference/src/source.rs):
``Source`` (:445-471), ``Seekable`` (:499-541), ``MemSource`` (:543-639), ``FrameBuf`` (:115-299)."""
from __future__ import annotations

import numpy as np

from .error import SourceError, VerifyError


class FrameBuf:
    """Planar multi-channel frame buffer: ``samples[ch * size + t]``, valid ``t < filled_size``."""

    def __init__(self, channels: int, size: int):
        if not 1 <= channels <= 8:
            raise VerifyError("FrameBuf::with_size (channels)", "must be in 1..=8")
        if not 32 <= size <= 32767:
            raise VerifyError("FrameBuf::with_size (block size)", "must be in 32..=32767")
        self.samples = np.zeros(channels * size, np.int32)
        self._size = size
        self._channels = channels
        self._filled = 0

    @staticmethod
    def with_size(channels: int, size: int) -> "FrameBuf":
        return FrameBuf(channels, size)

    def size(self) -> int:
        return self._size

    def filled_size(self) -> int:
        return self._filled

    def channels(self) -> int:
        return self._channels

    def fill_interleaved(self, interleaved) -> None:
        """Fill::fill_interleaved (src/source.rs:278-285): deinterleave into the planar layout."""
        x = np.asarray(interleaved, np.int32).reshape(-1, self._channels)
        if len(x) > self._size:
            raise SourceError("more samples than the buffer size")
        planar = self.samples.reshape(self._channels, self._size)
        planar[:, : len(x)] = x.T
        self._filled = len(x)

    def fill_le_bytes(self, data: bytes, bytes_per_sample: int) -> None:
        """Fill::fill_le_bytes (src/source.rs:287-299)."""
        b = np.frombuffer(data, np.uint8).reshape(-1, bytes_per_sample).astype(np.int32)
        v = np.zeros(len(b), np.int32)
        for i in range(bytes_per_sample):
            v |= b[:, i] << (8 * i)
        shift = 32 - 8 * bytes_per_sample
        v = (v << shift) >> shift
        self.fill_interleaved(v)

    def channel_slice(self, ch: int) -> np.ndarray:
        return self.samples[ch * self._size: ch * self._size + self._filled]


class MemSource:
    """Source backed by an in-memory interleaved sample array (src/source.rs:543-639)."""

    def __init__(self, samples, channels: int, bits_per_sample: int, sample_rate: int):
        self.samples = np.ascontiguousarray(samples, np.int32).reshape(-1)
        self._channels = channels
        self._bps = bits_per_sample
        self._rate = sample_rate
        self._pos = 0

    @staticmethod
    def from_samples(samples, channels: int, bits_per_sample: int, sample_rate: int) -> "MemSource":
        return MemSource(samples, channels, bits_per_sample, sample_rate)

    def channels(self) -> int:
        return self._channels

    def bits_per_sample(self) -> int:
        return self._bps

    def sample_rate(self) -> int:
        return self._rate

    def len_hint(self):
        return len(self.samples) // self._channels

    def __len__(self) -> int:
        return len(self.samples) // self._channels

    def read_samples(self, block_size: int, dest: FrameBuf) -> int:
        """Source::read_samples (src/source.rs:456-466): returns samples per channel read, 0 at EOF."""
        a = self._pos * self._channels
        b = min(len(self.samples), (self._pos + block_size) * self._channels)
        dest.fill_interleaved(self.samples[a:b])
        n = (b - a) // self._channels
        self._pos += n
        return n

    def read_samples_from(self, offset: int, block_size: int, dest: FrameBuf) -> int:
        """Seekable::read_samples_from (src/source.rs:512-521)."""
        self._pos = offset
        return self.read_samples(block_size, dest)

    def as_interleaved(self) -> np.ndarray:
        return self.samples.reshape(-1, self._channels)
