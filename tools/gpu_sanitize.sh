# compute-sanitizer (memcheck, then racecheck) over a few small encodes through the C ABI
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flacenc_rs_b200 import sigen
from flacenc_rs_b200.config import Encoder
from flacenc_rs_b200.encoder import Context
from conftest import pack_pcm
for (ch, bps, cont, block, frames, tail) in [(2, 16, 2, 4096, 9, 2728), (2, 24, 3, 4608, 3, 100), (8, 24, 3, 1024, 5, 37), (3, 16, 2, 128, 40, 5), (1, 16, 2, 4096, 33, 0)]:
    n = block * frames + tail
    chans = [sigen.Sine(23 + 5 * c, 0.6).noise(0.02, seed=40 + c).to_vec_quantized(bps, n) for c in range(ch)]
    x = np.stack(chans, axis=1)
    with Context(Encoder(block_size=block).into_verified(), ch, bps, 44100, block) as ctx:
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, cont), cont, n)
        print(ch, bps, block, len(sizes), int(sum(sizes)), ctx.timing().fused_frames, ctx.timing().fallback_frames)
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py 2>&1 | grep -v "^=========  *$" | tail -12
done
