"""Analysis-kernel time of K1 (thread per variant) against K1S (CTA per variant) over launch sizes, CD stereo frames."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from flacenc_rs_b200 import sigen  # noqa: E402
from flacenc_rs_b200.config import Encoder  # noqa: E402
from flacenc_rs_b200.encoder import Context, pack_samples  # noqa: E402

x_all = sigen.noisy_sine_pcm(4096 * 1200, 2, 16, 44100, config_id=2)
for frames in (54, 108, 222, 333, 444, 555, 666, 888, 1110):
    n = frames * 4096
    packed = pack_samples(x_all[:n], 2)
    row = []
    for small in ("0", "1"):
        os.environ["FB200_K1_SMALL"] = small
        with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
            cap = frames * ctx.max_frame_bytes()
            d_in = torch.from_numpy(packed).cuda()
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            sizes = np.zeros(frames, np.uint32)
            ts = []
            for i in range(6):
                ctx.encode_device(d_in.data_ptr(), 2, n, d_out.data_ptr(), cap, 0, sizes)
                t = ctx.timing()
                if i >= 2:
                    ts.append((t.k_analyze_ms, t.total_ms))
            row.append((round(float(np.mean([a for a, _ in ts])), 4), round(float(np.mean([b for _, b in ts])), 4)))
    print(f"frames {frames:5d} variants {frames * 4:5d}  K1 analyze/total ms {row[0]}  K1S {row[1]}", flush=True)
