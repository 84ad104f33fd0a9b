"""Encodes the C1-sized batch (108 frames of CD stereo) a few times: target for ncu captures of the small-batch case."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from conftest import pack_pcm
from flacenc_rs_b200 import sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 441000
x = sigen.noisy_sine_pcm(n, 2, 16, 44100)
pcm = pack_pcm(x, 2)
with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
    for _ in range(4):
        ctx.encode_interleaved(pcm, 2, n)
        t = ctx.timing()
    print("analyze ms", t.k_analyze_ms, "total", t.total_ms)
