# one-off (round 2): direct-MSE / IRLS tests, K1C / K1D / K1I timing at 2 min of CD stereo, then the ncu capture r2e
python -m pytest tests -m gpu -x -q -k "direct_mse or irls" 2>&1 | tail -2
python - <<'PY'
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from flacenc_rs_b200 import sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import Context, pack_samples
def run(tag, steps, env=None):
    if env: os.environ.update(env)
    e = Encoder(); e.subframe_coding.qlpc.use_direct_mse = True; e.subframe_coding.qlpc.window.type = "Rectangle"
    e.subframe_coding.qlpc.mae_optimization_steps = steps
    n = 120 * 44100
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=3); packed = pack_samples(x, 2)
    with Context(e.into_verified(), 2, 16, 44100, 4096) as ctx:
        nf = (n + 4095) // 4096; cap = nf * ctx.max_frame_bytes()
        d_in = torch.from_numpy(packed).cuda(); d_out = torch.empty(cap, dtype=torch.uint8, device="cuda"); sizes = np.zeros(nf, np.uint32)
        for _ in range(4):
            ctx.encode_device(d_in.data_ptr(), 2, n, d_out.data_ptr(), cap, 0, sizes); t = ctx.timing()
        print(tag, "analyze ms", round(t.k_analyze_ms, 4), "total", round(t.total_ms, 4), flush=True)
    if env:
        for k in env: os.environ.pop(k)
run("irls2", 2)
run("k1c (FB200_K1D=0)", 0, {"FB200_K1D": "0"})
run("k1d", 0)
PY
bash tools/gpu_profile2.sh r2e
