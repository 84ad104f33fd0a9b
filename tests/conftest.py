import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def load_fixture(name: str, ch: int) -> np.ndarray:
    """The reference's LE16 mono clips (src/test_helper.rs:81-116), copied to tests/golden/."""
    path = os.path.join(GOLDEN, f"testsignal.{name}.ch{ch}.bin")
    return np.fromfile(path, dtype="<i2").astype(np.int32)


def pack_pcm(signal: np.ndarray, container_bytes: int) -> np.ndarray:
    """Interleaved int32 samples -> packed little-endian bytes (what Fill::fill_le_bytes receives,
    /root/reference/src/source.rs:287-299); container 4 = int32 (fill_interleaved)."""
    x = np.ascontiguousarray(signal, np.int32).reshape(-1)
    if container_bytes == 4:
        return x.astype("<i4").view(np.uint8)
    if container_bytes == 1:
        return x.astype(np.int8).view(np.uint8)
    if container_bytes == 2:
        return x.astype("<i2").view(np.uint8)
    if container_bytes == 3:
        b = x.astype("<i4").view(np.uint8).reshape(-1, 4)
        return np.ascontiguousarray(b[:, :3]).reshape(-1)
    raise ValueError(container_bytes)


def crafted_huge_residual_stereo(order: int = 24) -> np.ndarray:
    """4096 x 2 samples of 24-bit stereo whose side channel drives the LPC residual above 2^26 (zigzag >= 2^27):
    six near-Nyquist tones give a high-gain predictor; the last `order` samples (where the Tukey window is ~0, so
    they barely move the coefficients) are full scale with the signs of the coefficients.  R = -L doubles it."""
    from oracle import oracle as O
    t = np.arange(4096)
    x = np.zeros(4096)
    for k in range(6):
        x += np.cos((np.pi - 0.03 - 0.05 * k) * t + 0.3 * k)
    s = (x / np.abs(x).max() * 8000000).astype(np.int32)
    for _ in range(3):
        coefs, _corr = O.lpc_from_autocorr(s, 1, 0.4, order)
        q, o, _shift = O.quantize_parameters(coefs, 15)
        for j in range(o):
            s[len(s) - 2 - j] = 8388607 if q[j] > 0 else -8388607
        s[len(s) - 1] = -8388607
    return np.stack([s, -s], axis=1)
