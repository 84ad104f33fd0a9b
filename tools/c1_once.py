"""One small-batch (BASELINE config 1: 10 s of CD stereo, 108 frames) device-resident encode, for ncu launch lists."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from flacenc_rs_b200 import sigen  # noqa: E402
from flacenc_rs_b200.config import Encoder  # noqa: E402
from flacenc_rs_b200.encoder import Context, pack_samples  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 441000
x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=1)
packed = pack_samples(x, 2)
with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
    n_frames = (n + 4095) // 4096
    cap = n_frames * ctx.max_frame_bytes()
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sizes = np.zeros(n_frames, np.uint32)
    for _ in range(reps):
        olen, _ = ctx.encode_device(d_in.data_ptr(), 2, n, d_out.data_ptr(), cap, 0, sizes)
        t = ctx.timing()
    print(olen, t.total_ms, t.k_analyze_ms, t.k_rice_ms, t.k_pack_ms, t.k_gather_ms, t.launches)
