"""Frame-range sharding over GPUs (what replaces par.rs's worker pool, /root/reference/src/par.rs:355-449).

Frames are independent, so GPU ``g`` of ``G`` encodes the contiguous range
``[ceil(F*g/G), ceil(F*(g+1)/G))`` and the host concatenates the ranges in order.
The only per-shard state STREAMINFO needs is min/max frame size, byte and frame counts
(component::StreamInfo::update_frame_info, src/component/datatype.rs:514-523).  No collective is involved on
the data path; the same arithmetic is done in C++ by fb200_encode_stream (csrc/fb_api.cu)."""
# The library's own multi-device call (fb200_encode_interleaved_sharded, csrc/fb_api.cu) shares a batch at CHUNK
# granularity: the batch is cut by ``chunk_schedule`` and chunk ``c`` is encoded on device ``c mod N`` (frame ranges
# again, just finer, so that every chunk's bytes can be copied straight to their final offset as soon as the chunk before
# it is done).  ``chunk_schedule`` / ``device_of_chunk`` mirror that arithmetic; tests compare them with the C++ code.
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import numpy as np


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    f0 = (n_frames * rank + world - 1) // world
    f1 = (n_frames * (rank + 1) + world - 1) // world
    return f0, f1


def sample_range(n_samples: int, block_size: int, f0: int, f1: int) -> Tuple[int, int]:
    """(first sample, sample count) per channel of frames [f0, f1); the stream's last frame may be short."""
    s0 = f0 * block_size
    return s0, max(0, min(n_samples - s0, (f1 - f0) * block_size))


def shard_stats(frame_sizes: np.ndarray) -> Dict[str, int]:
    if len(frame_sizes) == 0:
        return {"min_frame": 0xFFFFFFFF, "max_frame": 0, "bytes": 0, "frames": 0}
    return {"min_frame": int(frame_sizes.min()), "max_frame": int(frame_sizes.max()),
            "bytes": int(frame_sizes.astype(np.uint64).sum()), "frames": int(len(frame_sizes))}


def merge_shard_stats(stats: Iterable[Dict[str, int]]) -> Dict[str, int]:
    out = {"min_frame": 0xFFFFFFFF, "max_frame": 0, "bytes": 0, "frames": 0}
    for s in stats:
        out["min_frame"] = min(out["min_frame"], s["min_frame"])
        out["max_frame"] = max(out["max_frame"], s["max_frame"])
        out["bytes"] += s["bytes"]
        out["frames"] += s["frames"]
    return out


def chunk_schedule(total_frames: int, chunk_frames: int):
    """[(first frame, frames)] of the pipelined host path: the frames before a short final chunk (a quarter of the nominal
    size) spread evenly over chunks of at most the nominal size (fb_chunk_schedule in csrc/fb_api.cu)."""
    if total_frames == 0:
        return []
    chunk_frames = max(chunk_frames, 1)
    last = max(1, min(chunk_frames // 4, total_frames // 2))
    body = total_frames - last
    nb = (body + chunk_frames - 1) // chunk_frames
    out, f = [], 0
    for i in range(nb):
        take = body // nb + (1 if i < body % nb else 0)
        out.append((f, take))
        f += take
    out.append((f, last))
    return out


def device_of_chunk(c: int, n_devices: int) -> int:
    return c % n_devices
