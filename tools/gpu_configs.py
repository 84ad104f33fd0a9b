#!/usr/bin/env python
"""Throughput and parity spot-check of the other BASELINE.json configurations (C1, C3, C4-shape, C5 slice) on one GPU.
Prints one JSON line per configuration: device-resident and host (pipelined) throughput, path taken, stream ratio;
the first frames are compared byte for byte with the oracle."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flacenc_rs_b200 import sigen  # noqa: E402
from flacenc_rs_b200.config import Encoder  # noqa: E402
from flacenc_rs_b200.encoder import Context, pack_samples  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    # name, channels, bps, rate, block, seconds, config kwargs (oracle-style)
    ("C1 10 s CD stereo", 2, 16, 44100, 4096, 10, {}),
    ("C3 96k/24-bit stereo, block 4608, lpc_order 24", 2, 24, 96000, 4608, 120, {"lpc_order": 24}),
    ("C4-shape CD stereo, rectangle window", 2, 16, 44100, 4096, 600, {"window_type": 0}),
    ("C5 slice 48k/24-bit 8 ch", 8, 24, 48000, 4096, 120, {}),
    ("mono 16-bit 48k", 1, 16, 48000, 4096, 600, {}),
]


def make_cfg(kw):
    e = Encoder()
    if "lpc_order" in kw:
        e.subframe_coding.qlpc.lpc_order = kw["lpc_order"]
    if kw.get("window_type") == 0:
        e.subframe_coding.qlpc.window.type = "Rectangle"
    return e.into_verified()


ONLY = os.environ.get("FB_CFG_ONLY", "")        # run only the cases whose name starts with this
SCALE = int(os.environ.get("FB_CFG_SCALE", "1"))  # multiply the audio length
for name, ch, bps, rate, block, secs, kw in CASES:
    if not name.startswith(ONLY):
        continue
    n = secs * rate * SCALE
    x = sigen.noisy_sine_pcm(n, ch, bps, rate, config_id=3)
    cb = (bps + 7) // 8
    packed = pack_samples(x, cb)
    n_frames = (n + block - 1) // block
    with Context(make_cfg(kw), ch, bps, rate, block) as ctx:
        cap = n_frames * ctx.max_frame_bytes()
        d_in = torch.from_numpy(packed).cuda()
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        h_in = torch.empty(packed.nbytes, dtype=torch.uint8, pin_memory=True)
        h_in.numpy()[:] = packed
        h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        sizes = np.zeros(n_frames, np.uint32)
        for _ in range(2):
            olen, _ = ctx.encode_device(d_in.data_ptr(), cb, n, d_out.data_ptr(), cap, 0, sizes)
        ms = []
        for _ in range(3):
            olen, _ = ctx.encode_device(d_in.data_ptr(), cb, n, d_out.data_ptr(), cap, 0, sizes)
            t = ctx.timing()
            ms.append(t.total_ms)
        kern = {"ingest": t.k_ingest_ms, "analyze": t.k_analyze_ms, "plan": t.k_rice_ms, "pack": t.k_pack_ms,
                "fallback+scan": t.k_gather_ms}
        fused, fb = t.fused_frames, t.fallback_frames
        got, hs, _ = ctx.encode_interleaved(h_in.numpy(), cb, n, 0, out=h_out.numpy())
        e2e = []
        for _ in range(3):
            got, hs, _ = ctx.encode_interleaved(h_in.numpy(), cb, n, 0, out=h_out.numpy())
            e2e.append(ctx.timing().total_ms)
        assert len(got) == olen
    # parity on the first frames (the whole batch is covered by the pytest suite at smaller sizes)
    k = min(n_frames, 24)
    ref, ref_sizes = O.encode_frames(O.default_config(**kw), x[: k * block], ch, bps, rate, block)
    ok = bytes(got[: len(ref)]) == ref and list(hs[:k]) == list(ref_sizes)
    print(json.dumps({"config": name, "frames": n_frames, "device_gsamples_per_s": n / (np.mean(ms) / 1e3) / 1e9,
                      "device_ms": float(np.mean(ms)), "e2e_gsamples_per_s": n / (np.mean(e2e) / 1e3) / 1e9,
                      "kernel_ms": {a: round(b, 3) for a, b in kern.items()}, "fused_frames": int(fused),
                      "fallback_frames": int(fb), "ratio": olen / packed.nbytes, "first_frames_match_oracle": bool(ok)}),
          flush=True)
