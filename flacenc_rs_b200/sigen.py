"""Deterministic test-signal generators in the style of the reference's ``sigen`` module.

Mirrors /root/reference/src/sigen.rs (Dc :111-129, Sine :131-166, Square :168-190, Noise :192-219,
Mix :221-257, Clip :259-286, Switch :288-316, ``to_vec_quantized`` :35-53).  The reference's
``Noise`` draws from rand 0.8 ``StdRng`` (ChaCha12) which is not reproducible here; ours uses
numpy's Philox with the seeds documented in SURVEY.md section 8(d).  These are measurement and
test *inputs*; they are not part of the encode path.
"""
from __future__ import annotations

import numpy as np

PI32 = np.float32(np.pi)


class Signal:
    def fill(self, offset: int, n: int) -> np.ndarray:  # float32[n]
        raise NotImplementedError

    def to_vec_quantized(self, bits_per_sample: int, n: int, offset: int = 0) -> np.ndarray:
        """round(2^(bps-1) * x) clamped to the signed range (src/sigen.rs:35-53)."""
        assert 4 < bits_per_sample <= 24
        scale = np.float32(1 << (bits_per_sample - 1))
        lo = -(1 << (bits_per_sample - 1))
        hi = (1 << (bits_per_sample - 1)) - 1
        x = self.fill(offset, n).astype(np.float32) * scale
        # f32::round is half away from zero
        r = np.where(x >= 0, np.floor(x + np.float32(0.5)), np.ceil(x - np.float32(0.5)))
        return np.clip(r, lo, hi).astype(np.int32)

    def noise(self, amplitude: float, seed: int = 0) -> "Mix":
        return Mix(1.0, self, 1.0, Noise(seed, amplitude))

    def mix(self, other: "Signal") -> "Mix":
        return Mix(1.0, self, 1.0, other)

    def concat(self, offset_time: int, other: "Signal") -> "Switch":
        return Switch(self, offset_time, other)

    def clip(self) -> "Clip":
        return Clip(self)


class Dc(Signal):
    def __init__(self, offset: float):
        self.offset = np.float32(offset)

    def fill(self, offset, n):
        return np.full(n, self.offset, np.float32)


class Sine(Signal):
    def __init__(self, period: int, amplitude: float, initial_phase: float = 0.0):
        self.period = period
        self.amplitude = np.float32(amplitude)
        self.initial_phase = np.float32(initial_phase)

    def fill(self, offset, n):
        t = (np.arange(n, dtype=np.int64) + offset).astype(np.float32)
        ph = self.initial_phase + np.float32(2.0) * PI32 * t / np.float32(self.period)
        return (self.amplitude * np.sin(ph, dtype=np.float32)).astype(np.float32)


class Square(Signal):
    def __init__(self, period: int, amplitude: float):
        self.period = period
        self.amplitude = np.float32(amplitude)

    def fill(self, offset, n):
        t = np.arange(n, dtype=np.int64) + offset
        return np.where((t // self.period) % 2 == 0, self.amplitude, -self.amplitude).astype(np.float32)


class Noise(Signal):
    """Uniform noise in (-amplitude, amplitude); Philox keyed by (seed + offset)."""

    def __init__(self, seed: int, amplitude: float):
        self.seed = seed
        self.amplitude = np.float32(amplitude)

    def fill(self, offset, n):
        rng = np.random.Generator(np.random.Philox(key=(self.seed + offset) & ((1 << 64) - 1)))
        u = rng.random(n, dtype=np.float32)
        return (self.amplitude * np.float32(2.0) * (u - np.float32(0.5))).astype(np.float32)


class Mix(Signal):
    def __init__(self, w1: float, s1: Signal, w2: float, s2: Signal):
        self.w1, self.s1, self.w2, self.s2 = np.float32(w1), s1, np.float32(w2), s2

    def fill(self, offset, n):
        return (self.w1 * self.s1.fill(offset, n) + self.w2 * self.s2.fill(offset, n)).astype(np.float32)


class Clip(Signal):
    def __init__(self, inner: Signal):
        self.inner = inner

    def fill(self, offset, n):
        return np.clip(self.inner.fill(offset, n), np.float32(-1.0), np.float32(1.0))


class Switch(Signal):
    def __init__(self, s1: Signal, offset_time: int, s2: Signal):
        self.s1, self.offset_time, self.s2 = s1, offset_time, s2

    def fill(self, offset, n):
        out = self.s1.fill(offset, n)
        if n > self.offset_time:
            out[self.offset_time:] = self.s2.fill(offset + self.offset_time, n - self.offset_time)
        return out


def noisy_sine_pcm(n: int, channels: int, bits_per_sample: int, sample_rate: int, config_id: int = 1,
                   chunk: int = 1 << 22) -> np.ndarray:
    """The benchmark input of SURVEY.md section 8(d): per channel c, a 440 Hz sine (amp 0.8, phase
    0.37*c rad) under a slow 0.1 Hz amplitude ramp, mixed 1:1 with uniform noise (amp 0.2),
    quantised like ``to_vec_quantized``.  Seeds 0xF1AC0000 + config_id*16 + channel.
    Returns an interleaved int32 array of shape (n, channels)."""
    out = np.empty((n, channels), np.int32)
    scale = np.float32(1 << (bits_per_sample - 1))
    lo = -(1 << (bits_per_sample - 1))
    hi = (1 << (bits_per_sample - 1)) - 1
    for c in range(channels):
        rng = np.random.Generator(np.random.Philox(key=0xF1AC0000 + config_id * 16 + c))
        for s in range(0, n, chunk):
            m = min(chunk, n - s)
            t = np.arange(s, s + m, dtype=np.float64)
            ramp = 0.75 + 0.25 * np.sin(2.0 * np.pi * 0.1 * t / sample_rate + 0.5 * c)
            sine = 0.8 * ramp * np.sin(2.0 * np.pi * 440.0 * t / sample_rate + 0.37 * c)
            noise = 0.2 * 2.0 * (rng.random(m, dtype=np.float32).astype(np.float64) - 0.5)
            x = ((sine + noise).astype(np.float32)) * scale
            r = np.where(x >= 0, np.floor(x + np.float32(0.5)), np.ceil(x - np.float32(0.5)))
            out[s:s + m, c] = np.clip(r, lo, hi).astype(np.int32)
    return out
