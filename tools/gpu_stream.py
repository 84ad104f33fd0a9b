#!/usr/bin/env python
"""Stream-level calls timed in isolation (GPU box): single stream vs the MD5 floor, 16-stream batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flacenc_rs_b200 import _ffi, sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import encode_streams_with_fixed_block_size, encode_with_fixed_block_size, pack_samples
from flacenc_rs_b200.source import MemSource
RATE = 44100
ns = 600 * RATE
x = sigen.noisy_sine_pcm(ns, 2, 16, RATE, config_id=2)
src = MemSource.from_samples(x, 2, 16, RATE)
cfg = Encoder().into_verified()
encode_with_fixed_block_size(cfg, src, 4096)
for _ in range(4):
    t0 = time.perf_counter(); st = encode_with_fixed_block_size(cfg, src, 4096); dt = time.perf_counter() - t0
    packed = pack_samples(x, 2); dig = np.zeros(16, np.uint8)
    t0 = time.perf_counter(); _ffi.lib().fb200_md5(packed.ctypes.data, packed.nbytes, dig.ctypes.data); md = time.perf_counter() - t0
    print(f"stream {dt*1e3:.1f} ms  md5 floor {md*1e3:.1f} ms  frac {md/dt:.3f}")
seg = ns // 4
srcs = [MemSource.from_samples(x[(i % 4) * seg:(i % 4 + 1) * seg], 2, 16, RATE) for i in range(16)]
encode_streams_with_fixed_block_size(cfg, srcs[:4], 4096)
for _ in range(4):
    t0 = time.perf_counter(); outs = encode_streams_with_fixed_block_size(cfg, srcs, 4096); dt = time.perf_counter() - t0
    print(f"batch 16 x {seg/RATE:.0f} s: {dt*1e3:.1f} ms  {16*seg/dt/1e9:.2f} G samples/s")
