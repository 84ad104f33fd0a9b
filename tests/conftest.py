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
