# compute-sanitizer (memcheck, then racecheck) over small encodes through the C ABI; summaries go to gpurun_out/sanitize_<tag>.log
tag=${1:-r2}
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flacenc_rs_b200 import sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import Context, encode_interleaved_sharded
from conftest import pack_pcm, crafted_huge_residual_stereo
def run(ch, bps, cont, block, x, **kw):
    e = Encoder(block_size=block)
    if kw.get("lpc_order"): e.subframe_coding.qlpc.lpc_order = kw["lpc_order"]
    if kw.get("use_direct_mse"): e.subframe_coding.qlpc.use_direct_mse = True; e.subframe_coding.qlpc.window.type = "Rectangle"
    e.subframe_coding.qlpc.mae_optimization_steps = kw.get("mae_optimization_steps", 0)
    e.subframe_coding.qlpc.ext_order_search = kw.get("ext_lpc_order_search", 0)
    e.subframe_coding.qlpc.ext_precision_search = kw.get("ext_lpc_precision_search", 0)
    if kw.get("bitcount"):
        from flacenc_rs_b200.config import OrderSel
        e.subframe_coding.fixed.order_sel = OrderSel.BitCount()
    n = len(x)
    with Context(e.into_verified(), ch, bps, 44100, block) as ctx:
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, cont), cont, n)
        t = ctx.timing()
        print(ch, bps, cont, block, kw, len(sizes), int(sum(sizes)), "fused", t.fused_frames, "fallback", t.fallback_frames, flush=True)
def sig(ch, bps, n, seed=40):
    return np.stack([sigen.Sine(23 + 5 * c, 0.6).noise(0.02, seed=seed + c).to_vec_quantized(bps, n) for c in range(ch)], axis=1)
# 16-bit stereo (PCM pairs path) with an ODD tail frame (2728 = 8 x 341), 24-bit stereo order 24, 8 ch 24-bit, 3 ch, mono
run(2, 16, 2, 4096, sig(2, 16, 4096 * 9 + 2728))
run(2, 24, 3, 4608, sig(2, 24, 4608 * 3 + 100), lpc_order=24)
run(8, 24, 3, 1024, sig(8, 24, 1024 * 5 + 37))
run(3, 16, 2, 128, sig(3, 16, 128 * 40 + 5))
run(1, 16, 2, 4096, sig(1, 16, 4096 * 33))
# 16-bit stereo in a 4-byte container (ingest kernel + planar store), odd block size (every frame on the ODD instances)
run(2, 16, 4, 4096, sig(2, 16, 4096 * 4 + 1000))
run(2, 16, 2, 1001, sig(2, 16, 1001 * 5 + 333))
# a frame the fused kernels hand to the generic kernels (residual >= 2^26), between two ordinary frames
big = crafted_huge_residual_stereo()
run(2, 24, 3, 4096, np.concatenate([sig(2, 24, 4096), big, sig(2, 24, 4096, 7)]), lpc_order=24)
# direct-MSE estimator: K1C (CTA per variant), then K1D and the thread-per-variant K1 forced on the same small inputs
run(2, 16, 2, 4096, sig(2, 16, 4096 * 3 + 100), use_direct_mse=1)
# ... and its IRLS-MAE refinement (K1I), frames of several staging tiles and a 24-bit 3-channel case
run(2, 16, 2, 4096, sig(2, 16, 4096 * 3 + 100), use_direct_mse=1, mae_optimization_steps=2)
run(3, 24, 3, 1500, sig(3, 24, 1500 * 2 + 77), use_direct_mse=1, mae_optimization_steps=1, lpc_order=24)
# the probing instances of the plan kernel: BitCount order selection, the LPC order search extension (with a fallback frame)
run(2, 16, 2, 4096, sig(2, 16, 4096 * 3 + 100), bitcount=1)
run(2, 16, 2, 4096, sig(2, 16, 4096 * 3 + 100), ext_lpc_order_search=4, ext_lpc_precision_search=3)
run(2, 24, 3, 4096, np.concatenate([sig(2, 24, 4096), big]), lpc_order=24, ext_lpc_order_search=3, bitcount=1)
os.environ["FB200_K1_SMALL"] = "0"
run(2, 16, 2, 4096, sig(2, 16, 4096 * 3 + 100), use_direct_mse=1)
run(2, 16, 2, 4096, sig(2, 16, 4096 * 9 + 2728))
run(8, 24, 3, 1024, sig(8, 24, 1024 * 5 + 37))
del os.environ["FB200_K1_SMALL"]
# frame-range sharding over three contexts, chunks of 3 frames
os.environ["FB200_CHUNK_FRAMES"] = "3"
x = sig(2, 16, 1024 * 20 + 99)
ctxs = [Context(Encoder().into_verified(), 2, 16, 44100, 1024) for _ in range(3)]
got, sizes = encode_interleaved_sharded(ctxs, pack_pcm(x, 2), 2, len(x))
print("sharded", len(sizes), len(got), flush=True)
for c in ctxs: c.close()
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool  ($(git rev-parse --short HEAD 2>/dev/null || echo snapshot), $(date -u +%FT%TZ))"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py 2>&1 | grep -v "^=========  *$" | tail -25
done > gpurun_out/sanitize_${tag}.log 2>&1
tail -8 gpurun_out/sanitize_${tag}.log
