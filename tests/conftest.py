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


def shift_cases():
    """max|coef| values around every power of two 2^k: mantissa offsets m = 0..300 ulps above it, the ulps just below
    the next one, and random mantissas -- where ceil(log2(x)) read off the exponent and libm's rounded log2 can part
    ways (/root/reference/src/lpc.rs:234-255)."""
    rng = np.random.default_rng(7)
    vals = [0.0, 5e-324, 2.2250738585072014e-308, 1e-300, 1e300, float("inf")]
    for k in list(range(-40, 41)) + [-1000, -500, 500, 1000]:
        base = np.float64(2.0) ** k
        b0 = int(np.array([base], np.float64).view(np.uint64)[0])
        ms = list(range(0, 301)) + [(1 << 52) - d for d in range(1, 40)] + [int(v) for v in rng.integers(1, 1 << 52, 40)]
        ms += [1 << j for j in range(9, 52)]
        vals.extend(np.array([b0 + m for m in ms], np.uint64).view(np.float64).tolist())
    return vals


def fuzz_frame_case(rng):
    """One case in the shape of the reference's fuzz target (/root/reference/fuzz/fuzz_targets/frame_encode.rs:197-212):
    ONE frame of block size 32..32767, bits per sample 8..24, 1..8 channels, every toggle of the configuration random
    (orders 1..24, precision 1..15, Tukey alpha or Rectangle, max Rice parameter 0..30, fixed max order 0..4), the signal
    a random composition of DC, noise, sine, mixes and switches, optionally clipped.  Returns (signal, channels, bps,
    rate, block size, config kwargs)."""
    channels = int(rng.integers(1, 9))
    bps = int(rng.choice([8, 12, 16, 20, 24]))
    # block sizes: the whole range, with a bias to the extremes and to sizes whose Rice partitions are odd
    pick = rng.random()
    if pick < 0.25:
        block = int(rng.integers(32, 200))
    elif pick < 0.5:
        block = int(rng.integers(16000, 32768))
    elif pick < 0.6:
        block = int(rng.choice([32767, 32766, 32764, 16384, 8192, 4608, 4096, 1152, 576, 192]))
    else:
        block = int(rng.integers(200, 16000))
    full = float((1 << (bps - 1)) - 1)

    def signal(n, depth=0):
        kind = int(rng.integers(0, 5 if depth < 3 else 3))
        t = np.arange(n, dtype=np.float64)
        if kind == 0:
            x = np.full(n, rng.random())
        elif kind == 1:
            x = rng.random() * rng.uniform(-1, 1, n)
        elif kind == 2:
            period = max(1.0, n * rng.random())
            x = rng.random() * np.sin(2 * np.pi * t / period + 2 * np.pi * rng.random())
        elif kind == 3:
            a = rng.random()
            x = a * signal(n, depth + 1) + (1 - a) * signal(n, depth + 1)
        else:
            k = int(n * rng.random())
            x = np.concatenate([signal(k, depth + 1), signal(n - k, depth + 1)]) if 0 < k < n else signal(n, depth + 1)
        return x

    chans = []
    for _ in range(channels):
        x = signal(block) * (4.0 if rng.random() < 0.3 else 1.0)  # some channels overdrive and clip
        chans.append(np.clip(np.round(x * full), -full - 1, full).astype(np.int32))
    cfg = {
        "use_leftside": int(rng.random() < 0.7), "use_rightside": int(rng.random() < 0.7), "use_midside": int(rng.random() < 0.7),
        "use_constant": int(rng.random() < 0.8), "use_fixed": int(rng.random() < 0.8), "use_lpc": int(rng.random() < 0.8),
        "fixed_max_order": int(rng.integers(0, 5)), "prc_max_parameter": int(rng.integers(0, 31)),
        "lpc_order": int(rng.integers(1, 25)), "quant_precision": int(rng.integers(1, 16)),
        "fixed_order_sel": int(rng.random() < 0.8),
    }
    if rng.random() < 0.3:
        cfg["window_type"] = 0
    else:
        cfg["tukey_alpha"] = float(rng.random())
    if rng.random() < 0.15:
        cfg["use_direct_mse"] = 1
    rate = int(rng.choice([8000, 44100, 48000, 96000, 11025]))
    if cfg.get("use_direct_mse") and rng.random() < 0.5:  # (drawn last: earlier draws keep their values)
        cfg["mae_optimization_steps"] = int(rng.integers(1, 4))
    return np.stack(chans, axis=1), channels, bps, rate, block, cfg


def random_case(rng):
    """One seeded fuzz case: (signal, channels, bps, rate, block size, first frame number, oracle-style config kwargs)."""
    channels = int(rng.choice([1, 2, 2, 2, 3, 4, 6, 8]))
    bps = int(rng.choice([8, 12, 16, 16, 20, 24]))
    block = int(rng.choice([64, 96, 128, 192, 256, 500, 576, 1024, 1152, 2048, 2304, 4096, 4608, 1000, 3136]))
    frames = int(rng.integers(1, 4))
    n = block * (frames - 1) + int(rng.integers(1, block + 1))
    kind = int(rng.integers(0, 6))
    t = np.arange(n)
    full = (1 << (bps - 1)) - 1
    chans = []
    for c in range(channels):
        if kind == 0:      # tone + noise
            x = 0.6 * np.sin(t * (0.01 + 0.02 * rng.random()) + c) + 0.05 * rng.standard_normal(n)
        elif kind == 1:    # random walk
            x = np.cumsum(rng.standard_normal(n)) / 40.0
        elif kind == 2:    # loud / quiet halves
            x = np.where(t < n // 2, 0.8 * rng.uniform(-1, 1, n), 0.002 * rng.uniform(-1, 1, n))
        elif kind == 3:    # sparse impulses on silence
            x = np.zeros(n)
            x[rng.integers(0, n, max(1, n // 200))] = rng.uniform(-1, 1, max(1, n // 200))
        elif kind == 4:    # full-scale white noise (verbatim territory)
            x = rng.uniform(-1, 1, n)
        else:              # decaying chirp
            x = np.exp(-t / (n / 3.0)) * np.sin(t * t * 1e-5 + c)
        chans.append(np.clip(np.round(x * full), -full - 1, full).astype(np.int32))
    if channels == 2 and rng.random() < 0.3:
        chans[1] = chans[0] + rng.integers(-2, 3, n).astype(np.int32)  # near-identical channels: side-channel wins
        chans[1] = np.clip(chans[1], -full - 1, full)
    cfg = {}
    if rng.random() < 0.5:
        cfg["lpc_order"] = int(rng.integers(1, 25))
    if rng.random() < 0.3:
        cfg["quant_precision"] = int(rng.integers(2, 16))
    if rng.random() < 0.2:
        cfg["window_type"] = 0
    if rng.random() < 0.2:
        cfg["tukey_alpha"] = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    if rng.random() < 0.2:
        cfg["prc_max_parameter"] = int(rng.choice([2, 7, 14, 20, 30]))
    if rng.random() < 0.15:
        cfg["fixed_max_order"] = int(rng.integers(0, 5))
    if rng.random() < 0.15:
        cfg["approx_ent_partitions"] = int(rng.choice([1, 4, 8, 32, 64]))
    if rng.random() < 0.1:
        cfg["use_lpc"] = 0
    if rng.random() < 0.1:
        cfg["use_fixed"] = 0
    if rng.random() < 0.1:
        cfg["fixed_order_sel"] = 0
    rate = int(rng.choice([8000, 22050, 44100, 48000, 96000, 12345]))
    first = int(rng.choice([0, 1, 127, 128, 70000, (1 << 31) - 10]))
    if rng.random() < 0.2:  # (drawn last, so the cases above are the ones earlier versions of the suite ran)
        cfg["use_direct_mse"] = 1
        if rng.random() < 0.5:
            cfg["mae_optimization_steps"] = int(rng.integers(1, 4))
    return np.stack(chans, axis=1), channels, bps, rate, block, first, cfg


