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
