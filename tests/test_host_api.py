"""CPU-only tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/flacenc_b200.h declares, struct layouts agree with the header, host-side mirrors of the reference's
config / source / error modules behave like the reference, and the N>1 frame-range sharding logic is
consistent across ranks (world_size-2 gloo).  No compute call is made: there is no GPU here."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

from flacenc_rs_b200 import _ffi
from flacenc_rs_b200.config import Encoder, Fixed, OrderSel, Qlpc, SubFrameCoding, Verified, Window
from flacenc_rs_b200.error import EncodeError, SourceError, VerifyError, raise_for_code
from flacenc_rs_b200.source import FrameBuf, MemSource
from flacenc_rs_b200 import sharding

HEADER = os.path.join(ROOT, "include", "flacenc_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 15
    L = _ffi.lib()
    for n in names:
        assert hasattr(L, n), f"libflacenc_b200.so does not export {n}"
        assert n in _ffi.SYMBOLS, f"_ffi.SYMBOLS lacks {n}"
    assert sorted(_ffi.SYMBOLS) == names


def test_version_and_strerror_need_no_gpu():
    L = _ffi.lib()
    assert L.fb200_version().decode().startswith("flacenc_b200")
    assert L.fb200_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.fb200_strerror(3)


def test_integration_doc_declares_every_symbol_of_the_header():
    """INTEGRATION.md's Rust `extern "C"` block (the reference-side binding a maintainer would add) names every entry point
    include/flacenc_b200.h declares"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "flacenc_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    symbols = sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", header)))
    assert len(symbols) >= 20
    missing = [f for f in symbols if f"pub fn {f}(" not in doc]
    assert not missing, missing


def test_struct_sizes_match_header():
    # sizes computed by hand from include/flacenc_b200.h
    assert C.sizeof(_ffi.Config) == 21 * 4  # 19 fields of config::Encoder + 2 opt-in extensions
    assert C.sizeof(_ffi.SubframeInfo) == 8 * 4 + 64 + 256 + 8
    assert C.sizeof(_ffi.FrameInfo) == 16 + 8 * C.sizeof(_ffi.SubframeInfo)
    assert C.sizeof(_ffi.VariantTaps) == 25 * 8 + 24 * 8 + 64 + 4 * 4 + 5 * 8
    assert C.sizeof(_ffi.Timing) == 9 * 4 + 4 + 5 * 8


def test_no_device_means_error_not_fallback():
    """Without a usable device fb200_create must fail with FB200_ERR_CUDA (never encode on the CPU)."""
    L = _ffi.lib()
    if L.fb200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    cfg = _ffi.Config()
    L.fb200_config_default(C.byref(cfg))
    err = C.c_int(0)
    h = L.fb200_create(C.byref(cfg), 2, 16, 44100, 4096, 0, C.byref(err))
    assert not h and err.value == _ffi.ERR_CUDA


def test_config_default_matches_reference_defaults():
    """config::Encoder::default() (src/config.rs:97-107 and children)."""
    cfg = _ffi.Config()
    _ffi.lib().fb200_config_default(C.byref(cfg))
    pod = Encoder().to_pod()
    for name, _ in _ffi.Config._fields_:
        assert getattr(cfg, name) == pytest.approx(getattr(pod, name)), name
    assert (cfg.block_size, cfg.lpc_order, cfg.quant_precision, cfg.prc_max_parameter) == (4096, 10, 15, 30)
    assert (cfg.fixed_max_order, cfg.fixed_order_sel, cfg.approx_ent_partitions) == (4, 1, 16)
    assert cfg.window_type == 1 and cfg.tukey_alpha == pytest.approx(0.4)


@pytest.mark.parametrize("mutate,ok", [
    (lambda e: None, True),
    (lambda e: setattr(e.subframe_coding.qlpc, "lpc_order", 24), True),
    (lambda e: setattr(e.subframe_coding.qlpc, "lpc_order", 25), False),   # src/config.rs:304
    (lambda e: setattr(e.subframe_coding.qlpc, "lpc_order", 0), False),
    (lambda e: setattr(e.subframe_coding.qlpc, "quant_precision", 16), False),
    (lambda e: setattr(e.subframe_coding.qlpc, "use_direct_mse", True), True),   # the `experimental` covariance-method estimator is built
    (lambda e: setattr(e.subframe_coding.qlpc, "mae_optimization_steps", 2), True),   # ... and so is its IRLS-MAE refinement
    (lambda e: setattr(e.subframe_coding.qlpc, "mae_optimization_steps", -1), False),  # (a usize in the reference)
    (lambda e: setattr(e.subframe_coding.qlpc, "ext_order_search", 8), True),    # extension beyond the reference (opt-in)
    (lambda e: setattr(e.subframe_coding.qlpc, "ext_order_search", 9), False),
    (lambda e: setattr(e.subframe_coding.qlpc, "ext_precision_search", 4), True),
    (lambda e: setattr(e.subframe_coding.qlpc, "ext_precision_search", 5), False),
    (lambda e: (setattr(e.subframe_coding.qlpc, "ext_order_search", 6), setattr(e.subframe_coding.qlpc, "ext_precision_search", 3)), False),
    (lambda e: (setattr(e.subframe_coding.qlpc, "ext_order_search", 1), setattr(e.subframe_coding.qlpc, "use_direct_mse", True)), False),
    (lambda e: setattr(e.subframe_coding.qlpc, "window", Window.Tukey(1.5)), False),
    (lambda e: setattr(e.subframe_coding.qlpc, "window", Window.Rectangle()), True),
    (lambda e: setattr(e.subframe_coding.prc, "max_parameter", 31), False),
    (lambda e: setattr(e.subframe_coding.fixed, "max_order", 7), True),    # not rejected, SURVEY.md 9.2 #14
    (lambda e: setattr(e.subframe_coding.fixed, "order_sel", OrderSel.BitCount()), True),
    (lambda e: setattr(e, "block_size", 31), False),
    (lambda e: setattr(e, "block_size", 32768), False),
])
def test_config_verify(mutate, ok):
    e = Encoder()
    mutate(e)
    if ok:
        assert isinstance(e.into_verified(), Verified)
    else:
        with pytest.raises(VerifyError):
            e.into_verified()


def test_config_toml_form_of_the_reference():
    """config::Encoder's serde layout (src/config.rs:20-60; every table `serde(default)`): the module doc's example, the
    shapes of the reference's report/*.config.toml files, a dump / load round trip, and rejects"""
    doc = """
block_size = 4096
multithread = true

[stereo_coding]
use_leftside = true
use_rightside = true
use_midside = true

[subframe_coding]
use_constant = true
use_fixed = true
use_lpc = true

[subframe_coding.fixed]
max_order = 4

[subframe_coding.fixed.order_sel]
type = "ApproxEnt"
partitions = 32

[subframe_coding.qlpc]
lpc_order = 10
quant_precision = 15
use_direct_mse = false
mae_optimization_steps = 0

[subframe_coding.qlpc.window]
type = "Tukey"
alpha = 0.4

[subframe_coding.prc]
max_parameter = 14
"""
    e = Encoder.from_toml(doc)
    assert e.subframe_coding.fixed.order_sel == OrderSel.ApproxEnt(32) and e.subframe_coding.prc.max_parameter == 14
    assert Encoder.from_toml("") == Encoder()                                  # all defaults
    st = Encoder.from_toml("multithread = false\n")                           # report/st.config.toml's shape
    assert st.multithread is False and st.subframe_coding == Encoder().subframe_coding
    mt = Encoder.from_toml("multithread = true\nworkers = 1\n")               # report/mt1.config.toml's shape
    assert mt.workers == 1
    ex = Encoder.from_toml('[subframe_coding.qlpc]\nuse_direct_mse = true\n[subframe_coding.qlpc.window]\ntype = "Rectangle"\n')
    assert ex.subframe_coding.qlpc.use_direct_mse and ex.subframe_coding.qlpc.window == Window.Rectangle()
    bc = Encoder.from_toml('[subframe_coding.fixed.order_sel]\ntype = "BitCount"\n')
    assert bc.subframe_coding.fixed.order_sel.type == "BitCount"
    for cfg in (e, st, mt, ex, bc):
        assert Encoder.from_toml(cfg.to_toml()) == cfg
    x = Encoder()
    x.subframe_coding.qlpc.ext_order_search, x.subframe_coding.qlpc.ext_precision_search = 4, 2
    x.subframe_coding.qlpc.window = Window.Tukey(0.25)
    assert Encoder.from_toml(x.to_toml()) == x and "ext_order_search" not in Encoder().to_toml()
    for bad in ("block_size = \"big\"\n", "blocksize = 4096\n", "[subframe_coding.qlpc]\nlpc_order = true\n",
                '[subframe_coding.qlpc.window]\ntype = "Hann"\n', "[subframe_coding\n", '[subframe_coding.fixed.order_sel]\ntype = 3\n'):
        with pytest.raises(VerifyError):
            Encoder.from_toml(bad)
    # what loads still goes through Verify
    with pytest.raises(VerifyError):
        Encoder.from_toml("[subframe_coding.qlpc]\nlpc_order = 25\n").into_verified()


def test_error_code_mapping():
    raise_for_code(0)
    with pytest.raises(VerifyError):
        raise_for_code(1)
    with pytest.raises(SourceError):
        raise_for_code(2)
    with pytest.raises(EncodeError):
        raise_for_code(3)
    with pytest.raises(EncodeError):
        raise_for_code(4)


def test_framebuf_fill_le_bytes_sign_extension():
    """Fill::fill_le_bytes (src/source.rs:287-299): 3-byte samples are sign-extended."""
    fb = FrameBuf.with_size(2, 64)
    vals = np.array([0x7FFFFF, -0x800000, -1, 0, 12345, -54321], np.int32)
    raw = vals.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
    fb.fill_le_bytes(raw, 3)
    assert fb.filled_size() == 3
    assert list(fb.channel_slice(0)) == [0x7FFFFF, -1, 12345]
    assert list(fb.channel_slice(1)) == [-0x800000, 0, -54321]
    with pytest.raises(VerifyError):
        FrameBuf.with_size(9, 64)
    with pytest.raises(VerifyError):
        FrameBuf.with_size(2, 31)


def test_memsource_read_samples_and_eof():
    """Source::read_samples (src/source.rs:456-466): per-channel count, 0 at EOF, short last block."""
    x = np.arange(2 * 100, dtype=np.int32)
    src = MemSource.from_samples(x, 2, 16, 44100)
    fb = FrameBuf.with_size(2, 64)
    assert src.len_hint() == 100
    assert src.read_samples(64, fb) == 64
    assert list(fb.channel_slice(1)[:3]) == [1, 3, 5]
    assert src.read_samples(64, fb) == 36
    assert src.read_samples(64, fb) == 0
    assert src.read_samples_from(10, 8, fb) == 8 and fb.channel_slice(0)[0] == 20


# ---- frame-range sharding: GPU g of G gets frames [ceil(F*g/G), ceil(F*(g+1)/G)) ----

@pytest.mark.parametrize("n_frames,world", [(0, 2), (1, 2), (7, 2), (38760, 8), (108, 3), (5, 8), (421875, 4)])
def test_frame_ranges_partition_the_stream(n_frames, world):
    ranges = [sharding.frame_range(n_frames, g, world) for g in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_frames
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_sample_range_of_shard_handles_short_tail():
    n, bs = 441000, 4096
    F = (n + bs - 1) // bs
    total = 0
    for g in range(4):
        f0, f1 = sharding.frame_range(F, g, 4)
        s0, ns = sharding.sample_range(n, bs, f0, f1)
        assert s0 == f0 * bs
        total += ns
    assert total == n


def test_streaminfo_merge_matches_single_shard():
    sizes = np.array([900, 1200, 14, 4000, 77, 300, 650], np.uint32)
    whole = sharding.merge_shard_stats([sharding.shard_stats(sizes)])
    parts = sharding.merge_shard_stats([sharding.shard_stats(sizes[:3]), sharding.shard_stats(sizes[3:])])
    assert whole == parts == {"min_frame": 14, "max_frame": 4000, "bytes": int(sizes.sum()), "frames": 7}


@pytest.mark.parametrize("total,chunk", [(0, 2432), (1, 2432), (2, 1), (108, 2432), (5000, 2432), (38760, 2432), (38760, 807),
                                         (421875, 304), (62, 5), (7, 64)])
def test_chunk_schedule_partitions_the_batch_and_matches_the_library(total, chunk):
    """the chunk schedule of the (sharded) host path: chunks are contiguous, cover the batch, none larger than the nominal
    size; the Python mirror equals the C++ code (fb200_debug_chunk_schedule: host logic, no GPU needed)"""
    sched = sharding.chunk_schedule(total, chunk)
    assert sum(n for _, n in sched) == total
    f = 0
    for f0, n in sched:
        assert f0 == f and n >= 1 and n <= max(chunk, 1)
        f += n
    first = np.zeros(max(len(sched), 1) + 4, np.uint64)
    count = np.zeros(max(len(sched), 1) + 4, np.uint64)
    k = _ffi.lib().fb200_debug_chunk_schedule(total, chunk, first.ctypes.data, count.ctypes.data, len(first))
    assert k == len(sched)
    assert [(int(a), int(b)) for a, b in zip(first[:k], count[:k])] == sched
    # every device of a sharded call gets a fair share of the chunks
    for n_dev in (2, 3, 8):
        owners = [sharding.device_of_chunk(c, n_dev) for c in range(len(sched))]
        per = [owners.count(d) for d in range(n_dev)]
        assert max(per) - min(per) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from flacenc_rs_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, bs = 441000, 4096
F = (n + bs - 1) // bs
f0, f1 = sharding.frame_range(F, rank, world)
s0, ns = sharding.sample_range(n, bs, f0, f1)
# stand-in for the per-rank encode: frame i "encodes" to 100 + (i * 37) % 900 bytes
sizes = np.array([100 + (i * 37) % 900 for i in range(f0, f1)], np.uint32)
st = sharding.shard_stats(sizes)
# the only cross-rank traffic of the path: control-plane gather of per-shard stats + max-over-ranks of the time
gathered = [None] * world
dist.all_gather_object(gathered, (f0, f1, ns, st))
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# the library's own sharded call works at chunk granularity (chunk c on device c mod N): every rank derives the same
# schedule, the ranks' chunks are disjoint and cover the batch, and the byte offset of a chunk is the sum of the sizes
# of the chunks before it whichever rank encoded them
sched = sharding.chunk_schedule(F, 13)
mine = [(c, f0, nf) for c, (f0, nf) in enumerate(sched) if sharding.device_of_chunk(c, world) == rank]
mine_bytes = {c: int(sum(100 + (i * 37) % 900 for i in range(f0, f0 + nf))) for c, f0, nf in mine}
all_bytes = [None] * world
dist.all_gather_object(all_bytes, mine_bytes)
if rank == 0:
    merged_chunks = {}
    for d in all_bytes:
        assert not (set(d) & set(merged_chunks))
        merged_chunks.update(d)
    assert sorted(merged_chunks) == list(range(len(sched)))
    offs = np.concatenate([[0], np.cumsum([merged_chunks[c] for c in range(len(sched))])])
    assert offs[-1] == sum(100 + (i * 37) % 900 for i in range(F))
if rank == 0:
    gathered.sort()
    assert gathered[0][0] == 0 and gathered[-1][1] == F
    assert all(a[1] == b[0] for a, b in zip(gathered, gathered[1:]))
    assert sum(g[2] for g in gathered) == n
    merged = sharding.merge_shard_stats([g[3] for g in gathered])
    ref = sharding.shard_stats(np.array([100 + (i * 37) % 900 for i in range(F)], np.uint32))
    assert merged == ref, (merged, ref)
    assert t.item() == float(world)
    print("SHARD_OK", world, F)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_sharding(tmp_path):
    """N>1 host logic: two ranks over gloo take disjoint frame ranges that cover the stream; STREAMINFO
    statistics merged from the shards equal the single-shard ones; timing is reduced with MAX."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "SHARD_OK 2 108" in res.stdout
