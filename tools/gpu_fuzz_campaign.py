"""One-off parity campaign on the GPU box: many more seeds of the two fuzz generators of the test suite than the suite
runs (every case on the three device paths of tests/test_gpu_parity._compare, byte-compared with the oracle and decoded
back).  Usage: python tools/gpu_fuzz_campaign.py <first seed> <count> [seconds]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import fuzz_frame_case, random_case  # noqa: E402
import test_gpu_parity as T  # noqa: E402

first, count = int(sys.argv[1]), int(sys.argv[2])
budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
# "pipe": only the multi-frame generator, every signal repeated five times and the host pipeline cut into chunks of two
# frames (FB200_CHUNK_FRAMES), so that the chunked H2D / kernels / D2H path with its rotating buffer sets sees every
# configuration too
PIPE = len(sys.argv) > 4 and sys.argv[4] == "pipe"
if PIPE:
    import os
    os.environ["FB200_CHUNK_FRAMES"] = "2"
t0 = time.time()
done = bad = 0
kinds = {}
for seed in range(first, first + count):
    if time.time() - t0 > budget:
        break
    rng = np.random.default_rng(seed)
    try:
        if seed % 2 == 0 and not PIPE:
            x, channels, bps, rate, block, cfg = fuzz_frame_case(rng)
            first_frame = 0
        else:
            x, channels, bps, rate, block, first_frame, cfg = random_case(rng)
            if PIPE:
                x = np.concatenate([x, x[::-1], x, -x[::-1] // 2, x[: len(x) // 3]], axis=0)
                first_frame = min(first_frame, (1 << 31) - 64)
        if seed % 5 == 0 and not cfg.get("use_direct_mse"):
            cfg["ext_lpc_order_search"] = int(rng.integers(0, 9))   # the opt-in extensions, against the oracle's statement
            cfg["ext_lpc_precision_search"] = int(rng.integers(0, min(4, 8 - cfg["ext_lpc_order_search"]) + 1))
        T._compare(x, channels, bps, rate, block, first_frame=first_frame, oracle_threads=8, **cfg)
        for k in cfg:
            kinds[k] = kinds.get(k, 0) + 1
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print("MISMATCH / ERROR seed", seed, type(exc).__name__, str(exc)[:300], flush=True)
    done += 1
print(f"seeds {first}..{first + done - 1}: {done} cases, {bad} failures, {time.time() - t0:.0f} s", flush=True)
print("configuration keys exercised:", dict(sorted(kinds.items())))
