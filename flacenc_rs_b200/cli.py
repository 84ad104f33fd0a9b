"""`python -m flacenc_rs_b200.cli encode -o OUT.flac SOURCE.wav [SOURCE2.wav ...]` -- the encode command of the
reference's binary (/root/reference/flacenc-bin/src/main.rs:204-269) on top of the B200 library: WAV in (the
`HoundSource` of flacenc-bin/src/source.rs:86-133: PCM samples as little-endian bytes, 8-bit samples re-centred), the
stream from `fb200_encode_stream` / `fb200_encode_streams`, FLAC out.  No TOML configuration (defaults, plus the flags
below), no decode command."""
from __future__ import annotations

import argparse
import sys
import time
import wave
from typing import List

import numpy as np

from .config import Encoder
from .encoder import encode_streams_with_fixed_block_size, encode_with_fixed_block_size
from .source import MemSource


def wav_source(path: str) -> MemSource:
    """HoundSource: integer PCM WAV -> samples (8-bit WAV is unsigned: re-centred like flacenc-bin/src/source.rs:116-121)."""
    with wave.open(path, "rb") as w:
        ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = np.frombuffer(w.readframes(n), np.uint8)
    if width == 1:
        x = raw.astype(np.int32) - 128
    elif width == 2:
        x = raw.view("<i2").astype(np.int32)
    elif width == 3:
        b = raw.reshape(-1, 3).astype(np.int32)
        x = b[:, 0] | (b[:, 1] << 8) | ((b[:, 2].astype(np.int8).astype(np.int32)) << 16)
    else:
        raise ValueError(f"unsupported sample width {width} bytes (FLAC streams here carry 8..24 bits per sample)")
    return MemSource.from_samples(x.reshape(-1, ch), ch, width * 8, rate)


def main(argv: List[str] | None = None) -> int:
    ap = argparse.ArgumentParser(prog="flacenc_rs_b200.cli")
    sub = ap.add_subparsers(dest="command", required=True)
    enc = sub.add_parser("encode", help="encode WAV file(s) to FLAC on the GPU")
    enc.add_argument("-o", "--output", required=True, help="output FLAC path (with several sources: a directory-less prefix, "
                                                           "outputs are PREFIX<i>.flac)")
    enc.add_argument("-c", "--config", default=None, help="encoder configuration in the reference's TOML form "
                                                          "(flacenc-bin --config, src/config.rs:20-60)")
    enc.add_argument("--dump-config", default=None, help="write the configuration used to this path (flacenc-bin --dump-config)")
    enc.add_argument("--block-size", type=int, default=None)
    enc.add_argument("--lpc-order", type=int, default=None)
    enc.add_argument("--direct-mse", action="store_true", help="the `experimental` covariance-method LPC estimator")
    enc.add_argument("--devices", default="0", help="comma-separated device ordinals")
    enc.add_argument("source", nargs="+")
    args = ap.parse_args(argv)

    if args.config:
        with open(args.config, "r", encoding="utf-8") as f:
            cfg = Encoder.from_toml(f.read())
    else:
        cfg = Encoder()
    if args.block_size is not None:
        cfg.block_size = args.block_size
    args.block_size = cfg.block_size
    if args.lpc_order is not None:
        cfg.subframe_coding.qlpc.lpc_order = args.lpc_order
    if args.direct_mse:
        cfg.subframe_coding.qlpc.use_direct_mse = True
    vcfg = cfg.into_verified()
    if args.dump_config:
        with open(args.dump_config, "w", encoding="utf-8") as f:
            f.write(cfg.to_toml())
    devices = [int(d) for d in args.devices.split(",")]
    srcs = [wav_source(p) for p in args.source]
    t0 = time.perf_counter()
    if len(srcs) == 1:
        streams = [encode_with_fixed_block_size(vcfg, srcs[0], args.block_size, devices=devices)]
        outs = [args.output]
    else:
        streams = encode_streams_with_fixed_block_size(vcfg, srcs, args.block_size, devices=devices)
        outs = [f"{args.output}{i}.flac" for i in range(len(srcs))]
    dt = time.perf_counter() - t0
    for st, path in zip(streams, outs):
        with open(path, "wb") as f:
            f.write(st.write())
    secs = sum(len(s) / s.sample_rate() for s in srcs)
    print(f"encoded {len(srcs)} file(s), {secs:.1f} s of audio in {dt:.3f} s (inverse RTF {secs / max(dt, 1e-9):.0f})", file=sys.stderr)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
