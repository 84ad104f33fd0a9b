# experiment: fused kernel stopped after the analysis phase (results invalid, timing only)
for stop in 0 1; do
FB200_KF_DEBUG_STOP=$stop python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from flacenc_rs_b200 import sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import Context, pack_samples
n = 3600 * 44100
x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=2)
packed = pack_samples(x, 2)
nf = (n + 4095) // 4096
with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
    cap = nf * ctx.max_frame_bytes()
    d_in = torch.from_numpy(packed).cuda(); d_out = torch.empty(cap, dtype=torch.uint8, device='cuda')
    sizes = np.zeros(nf, np.uint32)
    for i in range(4):
        try:
            ctx.encode_device(d_in.data_ptr(), 2, n, d_out.data_ptr(), cap, 0, sizes)
        except Exception as e:
            pass
        t = ctx.timing()
    print(os.environ['FB200_KF_DEBUG_STOP'], 'encode ms', t.k_rice_ms, 'analyze', t.k_analyze_ms)
PY
done
