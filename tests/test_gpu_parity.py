"""GPU parity tests: the CUDA path, called through the C ABI (libflacenc_b200.so via ctypes), against the
CPU oracle on the same seeded inputs.  Bar: frame bytes bit-exact; floats of the analysis tier within 1e-5
relative (they are in fact bit-identical); decoded PCM identical to the input (tier 1).
Nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import crafted_huge_residual_stereo, fuzz_frame_case, load_fixture, pack_pcm, random_case, shift_cases
from flacenc_rs_b200 import _ffi, sigen
from flacenc_rs_b200.config import Encoder, Fixed, OrderSel, Qlpc, StereoCoding, SubFrameCoding, Window, Prc
from flacenc_rs_b200.encoder import (Context, StreamInfo, encode_fixed_size_frame, encode_with_fixed_block_size)
from flacenc_rs_b200.error import VerifyError
from flacenc_rs_b200.source import FrameBuf, MemSource
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def make_config(**kw) -> Encoder:
    """flat oracle-style keywords -> config::Encoder mirror"""
    e = Encoder()
    s = e.subframe_coding
    for k, v in kw.items():
        if k in ("use_leftside", "use_rightside", "use_midside"):
            setattr(e.stereo_coding, k, bool(v))
        elif k in ("use_constant", "use_fixed", "use_lpc"):
            setattr(s, k, bool(v))
        elif k == "fixed_max_order":
            s.fixed.max_order = v
        elif k == "fixed_order_sel":
            s.fixed.order_sel.type = "BitCount" if v == 0 else "ApproxEnt"
        elif k == "approx_ent_partitions":
            s.fixed.order_sel.partitions = v
        elif k == "lpc_order":
            s.qlpc.lpc_order = v
        elif k == "quant_precision":
            s.qlpc.quant_precision = v
        elif k == "use_direct_mse":
            s.qlpc.use_direct_mse = bool(v)
        elif k == "mae_optimization_steps":
            s.qlpc.mae_optimization_steps = v
        elif k == "ext_lpc_order_search":
            s.qlpc.ext_order_search = v
        elif k == "ext_lpc_precision_search":
            s.qlpc.ext_precision_search = v
        elif k == "window_type":
            s.qlpc.window.type = "Rectangle" if v == 0 else "Tukey"
        elif k == "tukey_alpha":
            s.qlpc.window.alpha = v
        elif k == "prc_max_parameter":
            s.prc.max_parameter = v
        elif k == "block_size":
            e.block_size = v
        else:
            raise KeyError(k)
    return e


def _compare(signal, channels, bps, rate, block_size, container=None, first_frame=0, check_infos=False,
             oracle_threads=1, **cfgkw):
    ocfg = O.default_config(**cfgkw)
    vcfg = make_config(**cfgkw).into_verified()
    assert bytes(vcfg.pod) == bytes(ocfg)  # the POD the C ABI receives equals the oracle's config
    signal = np.ascontiguousarray(signal, np.int32).reshape(-1, channels)
    n = len(signal)
    container = container or (bps + 7) // 8
    ref, ref_sizes = O.encode_frames(ocfg, signal, channels, bps, rate, block_size, first_frame_number=first_frame,
                                     nthreads=oracle_threads)
    # both device paths: the fused per-frame kernel (default, when the batch is eligible) and the generic kernels
    # (the analysis runs as a warp per variant for small launches -- nearly every case here -- so the classic
    # thread-per-variant kernel is forced on in the other modes)
    modes = [("0", "1", None), ("1", "1", "0")]
    if channels == 2 and bps == 16 and container == 2:
        modes.insert(1, ("0", "0", "0"))  # 16-bit stereo: also with planes instead of PCM pairs in every kernel
    for force_generic, kp_pairs, k1_small in modes:
        os.environ["FB200_FORCE_GENERIC"] = force_generic
        os.environ["FB200_KP_PAIRS"] = kp_pairs
        if k1_small is not None:
            os.environ["FB200_K1_SMALL"] = k1_small
        try:
            with Context(vcfg, channels, bps, rate, block_size) as ctx:
                got, sizes, infos = ctx.encode_interleaved(pack_pcm(signal, container), container, n, first_frame,
                                                           want_infos=check_infos)
                got = got.tobytes()
                t = ctx.timing()
                assert t.launches >= 6 and t.kernels_ms > 0
                if force_generic == "1":
                    assert t.fused_frames == 0 and t.fallback_frames == 0
        finally:
            os.environ.pop("FB200_FORCE_GENERIC", None)
            os.environ.pop("FB200_KP_PAIRS", None)
            os.environ.pop("FB200_K1_SMALL", None)
        assert list(sizes) == list(ref_sizes), f"force_generic={force_generic}"
        if got != ref:
            off = 0
            for i, s in enumerate(ref_sizes):
                assert got[off:off + s] == ref[off:off + s], f"frame {i} differs (size {s}), generic={force_generic}"
                off += s
        assert got == ref
    out, nf = O.decode_frames(got, channels, bps)
    assert np.array_equal(out, signal)
    return infos


def test_cd_stereo_c1_full():
    """BASELINE config 1: 10 s, 44.1 kHz, 16-bit stereo, block 4096 (108 frames, tail 2728)"""
    x = sigen.noisy_sine_pcm(441000, 2, 16, 44100, config_id=1)
    infos = _compare(x, 2, 16, 44100, 4096, check_infos=True)
    assert infos[107].block_size == 2728 and infos[0].block_size == 4096


@pytest.mark.parametrize("name", ["sus109", "sus6", "ras22", "ras103"])
def test_fixture_clips(name):
    x = np.stack([load_fixture(name, 0), load_fixture(name, 1)], axis=1)
    _compare(x, 2, 16, 44100, 4096)
    _compare(x[:, 0], 1, 16, 44100, 4096)


def test_96k_24bit_order24_block4608_c3():
    x = sigen.noisy_sine_pcm(4608 * 20, 2, 24, 96000, config_id=3)
    _compare(x, 2, 24, 96000, 4608, lpc_order=24, quant_precision=15)


def test_8ch_24bit_c5_slice():
    x = sigen.noisy_sine_pcm(4096 * 6 + 1000, 8, 24, 48000, config_id=5)
    _compare(x, 8, 24, 48000, 4096)


def test_c3_1000_frames_byte_exact():
    """BASELINE config 3 at scale: 1 000 frames + a tail of 96 kHz / 24-bit stereo, block 4608, lpc_order 24 -- every
    frame byte-compared with the oracle, on all device paths"""
    x = sigen.noisy_sine_pcm(4608 * 1000 + 1234, 2, 24, 96000, config_id=3)
    _compare(x, 2, 24, 96000, 4608, oracle_threads=os.cpu_count() or 1, lpc_order=24, quant_precision=15)


def test_c5_1000_frames_byte_exact():
    """BASELINE config 5 at scale: 1 000 frames + a tail of 48 kHz / 24-bit 8-channel audio (8 000 subframes)"""
    x = sigen.noisy_sine_pcm(4096 * 1000 + 777, 8, 24, 48000, config_id=5)
    _compare(x, 8, 24, 48000, 4096, oracle_threads=os.cpu_count() or 1)


@pytest.mark.parametrize("block_size", [1001, 4097, 4098, 1002])
def test_stereo_16bit_odd_block_sizes_frame_starts_unaligned(block_size):
    """16-bit stereo in a 2-byte container with block sizes that are not multiples of 4: every other frame starts at an
    address that is not 16-byte aligned, so neither the ingest kernel's 16-byte loads nor the pack kernel's staging of
    the packed PCM pairs may assume alignment (5 frames + tail)"""
    n = block_size * 5 + 333
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=11)
    _compare(x, 2, 16, 44100, block_size)


@pytest.mark.parametrize("channels,block_size", [(1, 1001), (1, 4097), (1, 4098), (3, 4098), (3, 1001), (5, 999), (2, 1001), (7, 4097)])
def test_packed_24bit_odd_block_sizes_frame_starts_unaligned(channels, block_size):
    """packed 24-bit samples (3-byte container): frames whose first byte is not 4-byte aligned must not take the
    ingest kernel's aligned-word fast path"""
    n = block_size * 4 + 100
    x = sigen.noisy_sine_pcm(n, channels, 24, 48000, config_id=12)
    _compare(x, channels, 24, 48000, block_size, container=3)


def test_experimental_config_c4_direct_mse():
    """BASELINE config 4 = report/experimental.config.toml: use_direct_mse + Rectangle window (covariance-method LPC,
    src/lpc.rs:852-913) on CD stereo; plus Tukey windows, other orders, 24-bit, and degenerate signals whose
    covariance matrix is not positive definite (regularised diagonal)"""
    x = sigen.noisy_sine_pcm(4096 * 12 + 1500, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 4096, use_direct_mse=1, window_type=0)
    _compare(x[: 4096 * 3], 2, 16, 44100, 4096, use_direct_mse=1)
    _compare(x[: 1024 * 5 + 77], 2, 16, 44100, 1024, use_direct_mse=1, lpc_order=24, tukey_alpha=0.1)
    _compare(x[: 1024 * 5, 0], 1, 16, 44100, 1024, use_direct_mse=1, lpc_order=1, quant_precision=5)
    y = sigen.noisy_sine_pcm(4608 * 3 + 100, 2, 24, 96000, config_id=3)
    _compare(y, 2, 24, 96000, 4608, use_direct_mse=1, window_type=0, lpc_order=16)
    _compare(np.zeros((3000, 2), np.int32), 2, 16, 44100, 1024, use_direct_mse=1, window_type=0)
    c = np.full((3000, 1), 1000, np.int32)
    c[1500:] = -77
    _compare(c, 1, 16, 44100, 1024, use_direct_mse=1, window_type=0, use_constant=0)
    imp = np.zeros((4096, 1), np.int32)
    imp[1000] = 30000
    _compare(imp, 1, 16, 44100, 4096, use_direct_mse=1, window_type=0)
    _compare(sigen.noisy_sine_pcm(20000 * 2 + 5, 2, 16, 44100, config_id=8), 2, 16, 44100, 20000, use_direct_mse=1, window_type=0)


def test_experimental_irls_mae_refinement():
    """use_direct_mse + mae_optimization_steps > 0 (the IRLS-MAE refinement of src/lpc.rs:814-850, the estimator the
    reference's own report/experimental.config.toml-style configs can select): byte-equal frames over step counts,
    orders, windows, sample formats and degenerate signals; the winning coefficients bit-equal to the oracle's"""
    x = sigen.noisy_sine_pcm(4096 * 6 + 1500, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 4096, use_direct_mse=1, mae_optimization_steps=2, window_type=0)
    _compare(x[: 4096 * 3], 2, 16, 44100, 4096, use_direct_mse=1, mae_optimization_steps=1)
    _compare(x[: 1024 * 5 + 77], 2, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=4, lpc_order=24, tukey_alpha=0.1)
    _compare(x[: 1024 * 5, 0], 1, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=7, lpc_order=1, quant_precision=5)
    y = sigen.noisy_sine_pcm(4608 * 3 + 100, 2, 24, 96000, config_id=3)
    _compare(y, 2, 24, 96000, 4608, use_direct_mse=1, mae_optimization_steps=2, window_type=0, lpc_order=16)
    _compare(np.zeros((3000, 2), np.int32), 2, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=2, window_type=0)
    c = np.full((3000, 1), 1000, np.int32)
    c[1500:] = -77
    _compare(c, 1, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=3, window_type=0, use_constant=0)
    imp = np.zeros((4096, 1), np.int32)
    imp[1000] = 30000
    _compare(imp, 1, 16, 44100, 4096, use_direct_mse=1, mae_optimization_steps=2, window_type=0)
    _compare(sigen.noisy_sine_pcm(20000 * 2 + 5, 2, 16, 44100, config_id=8), 2, 16, 44100, 20000, use_direct_mse=1,
             mae_optimization_steps=2, window_type=0)
    z = sigen.noisy_sine_pcm(1024 * 4, 8, 24, 48000, config_id=5)
    _compare(z, 8, 24, 48000, 1024, use_direct_mse=1, mae_optimization_steps=2)
    # the steps are ignored unless the direct-MSE estimator is selected (src/coding.rs:337-351)
    _compare(x[: 4096 * 3], 2, 16, 44100, 4096, use_direct_mse=0, mae_optimization_steps=3)
    # float tier
    sig = x[:4096, 0]
    cfg = make_config(use_direct_mse=1, mae_optimization_steps=4, window_type=0).into_verified()
    with Context(cfg, 1, 16, 44100, 4096) as ctx:
        taps, nv = ctx.analyze(pack_pcm(sig.reshape(-1, 1), 2), 2, 4096)
    assert nv == 1
    coefs, sums = O.lpc_with_irls_mae(sig, 0, 0.0, 10, 4)
    plain, corr, _ = O.lpc_with_direct_mse(sig, 0, 0.0, 10)
    assert np.array_equal(np.array(taps[0].lpc[:10]), coefs) and np.array_equal(np.array(taps[0].autocorr[:11]), corr)
    assert not np.array_equal(coefs, plain) and sums.min() < sums[0]


def test_device_irls_weight_matches_libm_exhaustive():
    """The device build of the IRLS weight (src/lpc.rs:828; powf(-1.2) by fb_powf_pos) against the host glibc's for every
    non-negative raw-error bit pattern under one normalizer, slices under others and of the negative errors"""
    lib = _ffi.lib()
    threads = min(32, os.cpu_count() or 1)
    step = 1 << 26
    out = np.empty(step, np.uint32)

    def sweep(first, end, norm):
        while first < end:
            cnt = min(step, end - first)
            assert lib.fb200_debug_irls_weight(0, first, cnt, norm, out.ctypes.data) == 0
            want = O.irls_weight_bits(first, cnt, norm, threads)
            bad = np.flatnonzero(out[:cnt] != want)
            assert len(bad) == 0, (norm, hex(first + int(bad[0])), hex(int(out[bad[0]])), hex(int(want[bad[0]])))
            first += cnt

    sweep(0, 0x7F800000 + 4096, 100.0)
    for norm in (1.0, 3.0, 32767.0, 8388607.0, 0.0):
        sweep(0x3F000000, 0x4C000000, norm)
    sweep(0xBF000000, 0xC1000000, 1000.0)


def test_rectangle_window_c4_shape():
    x = sigen.noisy_sine_pcm(4096 * 10, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 4096, window_type=0)


@pytest.mark.parametrize("cfgkw", [
    dict(use_lpc=0), dict(use_fixed=0), dict(use_constant=0), dict(fixed_order_sel=0),
    dict(use_leftside=0, use_rightside=0), dict(use_midside=0), dict(use_leftside=0, use_rightside=0, use_midside=0),
    dict(lpc_order=1, quant_precision=3), dict(lpc_order=7, quant_precision=12, tukey_alpha=0.1),
    dict(fixed_max_order=0), dict(fixed_max_order=2, approx_ent_partitions=32), dict(approx_ent_partitions=1),
    dict(approx_ent_partitions=64), dict(prc_max_parameter=0), dict(prc_max_parameter=3), dict(prc_max_parameter=14),
    dict(fixed_max_order=9), dict(tukey_alpha=1.0), dict(tukey_alpha=0.0), dict(lpc_order=16), dict(lpc_order=20),
])
def test_config_variations(cfgkw):
    x = sigen.Sine(60, 0.5).noise(0.05, seed=2).to_vec_quantized(16, 2 * 3000).reshape(3000, 2)
    _compare(x, 2, 16, 44100, 1024, **cfgkw)


@pytest.mark.parametrize("block_size", [32, 63, 64, 65, 100, 192, 576, 1000, 1152, 4000, 4095, 4097, 16384])
def test_block_sizes(block_size):
    n = block_size * 2 + block_size // 3 + 1
    x = sigen.Sine(77, 0.6).noise(0.02, seed=block_size).to_vec_quantized(16, n)
    _compare(x, 1, 16, 32000, block_size)


def test_max_block_size_8ch_global_pack_path():
    x = sigen.noisy_sine_pcm(32767 + 500, 8, 24, 96000, config_id=7)
    _compare(x, 8, 24, 96000, 32767)


@pytest.mark.parametrize("channels", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("bps,container", [(8, 1), (12, 2), (16, 2), (20, 3), (24, 3), (24, 4), (16, 4)])
def test_channels_and_sample_formats(channels, bps, container):
    n = 2500
    chans = [sigen.Sine(30 + 7 * c, 0.7).noise(0.01, seed=c).to_vec_quantized(bps, n) for c in range(channels)]
    _compare(np.stack(chans, axis=1), channels, bps, 48000, 1024, container=container)


@pytest.mark.parametrize("channels", [1, 2, 3, 4, 5, 6, 7, 8])
def test_many_short_frames_every_channel_count(channels):
    """70 frames + a short tail of 128-sample blocks: the analysis kernel's warps (32 channel variants each) span many
    frames, frames straddle warps for 3, 5, 6 and 7 channels (more than 32 staged rows), and the last warp mixes frame
    lengths."""
    n = 128 * 70 + 37
    chans = [sigen.Sine(23 + 5 * c, 0.6).noise(0.02, seed=40 + c).to_vec_quantized(16, n) for c in range(channels)]
    _compare(np.stack(chans, axis=1), channels, 16, 44100, 128)


def test_special_signals():
    _compare(np.zeros((3000, 2), np.int32), 2, 16, 44100, 1024)
    c = np.full((3000, 2), -1234, np.int32)
    c[:, 1] = 77
    _compare(c, 2, 16, 44100, 1024)
    fs = np.tile(np.array([[32767, -32768], [-32768, 32767]], np.int32), (1500, 1))
    _compare(fs, 2, 16, 44100, 1024)
    rng = np.random.default_rng(5)
    _compare(rng.integers(-(1 << 23), 1 << 23, (3000, 2)).astype(np.int32), 2, 24, 96000, 1024)
    imp = np.zeros((4096, 1), np.int32)
    imp[1000] = 30000
    _compare(imp, 1, 16, 44100, 4096)
    loudquiet = np.concatenate([rng.integers(-30000, 30000, 2048), rng.integers(-3, 3, 2048)]).astype(np.int32)
    _compare(loudquiet.reshape(-1, 1), 1, 16, 44100, 4096)


def test_pathological_residuals():
    """residuals >= 2^27: chunked saturating Rice tables (src/rice.rs:75-98) and 32-bit wrap (src/lpc.rs:375-389)"""
    t = np.arange(4096)
    smooth = (np.sin(t / 300.0) * 8000000).astype(np.int32)
    smooth[3000:] = np.where(np.arange(1096) % 2 == 0, 8388607, -8388608)
    _compare(smooth.reshape(-1, 1), 1, 24, 96000, 4096, lpc_order=24)
    _compare(np.stack([smooth, smooth[::-1]], axis=1), 2, 24, 96000, 4096, lpc_order=12, prc_max_parameter=5)
    alt = np.where(np.arange(4096) % 2 == 0, 8388607, -8388608).astype(np.int32).reshape(-1, 1)
    _compare(alt, 1, 24, 96000, 4096, fixed_order_sel=0)


def test_fused_kernel_is_the_path_taken_and_falls_back_only_when_it_must():
    """Default config: every frame is encoded by the fused per-frame kernels, also a tail frame whose finest Rice
    partitions are not a multiple of 4 samples (2728 = 8 x 341: the kernels' ODD instances).  A residual >= 2^26
    (zigzag >= 2^27) or a saturated table minimum hands the frame to the generic kernels; results stay byte-identical."""
    vcfg = Encoder().into_verified()
    n = 4096 * 5 + 2728
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100)
    with Context(vcfg, 2, 16, 44100, 4096) as ctx:
        ctx.encode_interleaved(pack_pcm(x, 2), 2, n)
        t = ctx.timing()
        assert (t.fused_frames, t.fallback_frames) == (6, 0)
        ctx.encode_interleaved(pack_pcm(x[: 4096 * 5 + 2048], 2), 2, 4096 * 5 + 2048)
        t = ctx.timing()
        assert (t.fused_frames, t.fallback_frames) == (6, 0)
    y = crafted_huge_residual_stereo()
    _compare(y, 2, 24, 96000, 4096, lpc_order=24)
    with Context(make_config(lpc_order=24).into_verified(), 2, 24, 96000, 4096) as ctx:
        ctx.encode_interleaved(pack_pcm(np.concatenate([y, y // 4, y]), 3), 3, 3 * 4096)
        t = ctx.timing()
        assert (t.fused_frames, t.fallback_frames) == (1, 2)
    with Context(make_config(fixed_order_sel=0).into_verified(), 2, 16, 44100, 4096) as ctx:
        ctx.encode_interleaved(pack_pcm(x, 2), 2, n)
        t = ctx.timing()
        assert (t.fused_frames, t.fallback_frames) == (6, 0)  # BitCount selection runs on the fused kernels too


def test_bitcount_order_selection_on_the_fused_path():
    """OrderSel::BitCount (src/coding.rs:241-262) in the plan kernel: signals whose best fixed order differs, max orders
    below 4, no LPC, a constant channel, RICE2 parameters, 24-bit order 24, 8 channels"""
    t = np.arange(8192, dtype=np.float64)
    rng = np.random.default_rng(21)
    sigs = [
        rng.integers(-20000, 20000, 8192),
        np.cumsum(rng.integers(-300, 300, 8192)),
        np.cumsum(np.cumsum(rng.integers(-3, 4, 8192))) % 30000,
        (20000 * np.sin(t / 40.0)).astype(np.int64) + rng.integers(-2, 3, 8192),
        (3000 * np.sin(t / 9.0) + 9000 * np.sin(t / 100.0)).astype(np.int64),
    ]
    for s in sigs:
        x = np.stack([s, s[::-1]], axis=1).astype(np.int32)
        _compare(x, 2, 16, 44100, 4096, fixed_order_sel=0)
        _compare(x[:, :1], 1, 16, 44100, 2048, fixed_order_sel=0, use_lpc=0)
    for mo in (0, 1, 2, 3):
        _compare(np.stack([sigs[3], sigs[4]], axis=1).astype(np.int32), 2, 16, 44100, 4096, fixed_order_sel=0, fixed_max_order=mo)
    _compare(np.stack([sigs[2], np.full(8192, 77)], axis=1).astype(np.int32), 2, 16, 44100, 1000, fixed_order_sel=0)
    _compare(rng.integers(-(1 << 22), 1 << 22, (4096, 2)).astype(np.int32), 2, 24, 96000, 1024, fixed_order_sel=0)
    y = sigen.noisy_sine_pcm(4608 * 3 + 100, 2, 24, 96000, config_id=3)
    _compare(y, 2, 24, 96000, 4608, fixed_order_sel=0, lpc_order=24)
    z = sigen.noisy_sine_pcm(1024 * 4 + 9, 8, 24, 48000, config_id=5)
    _compare(z, 8, 24, 48000, 1024, fixed_order_sel=0)
    with Context(make_config(fixed_order_sel=0).into_verified(), 2, 24, 96000, 4608) as ctx:
        ctx.encode_interleaved(pack_pcm(y, 3), 3, len(y))
        tm = ctx.timing()
        assert (tm.fused_frames, tm.fallback_frames) == (4, 0)


def test_ext_lpc_order_search_extension():
    """EXTENSION beyond the reference (config.ext_lpc_order_search, opt-in): lower LPC orders from the same
    autocorrelation, fewest bits wins.  The kernels (fused probing instances, generic, both analysis kernels) are
    byte-equal to the oracle's statement of it; streams decode to the input and are never larger."""
    rng = np.random.default_rng(33)
    n = 4096 * 5 + 700
    t = np.arange(n, dtype=np.float64)
    ar2 = np.zeros(n)
    e = rng.normal(0, 300, n)
    for i in range(2, n):
        ar2[i] = 1.027 * ar2[i - 1] - 0.9025 * ar2[i - 2] + 1.5 * e[i]
    sigs = [
        np.stack([ar2, np.roll(ar2, 7)], axis=1),
        np.stack([8000 * np.sin(t / 15.0) + rng.normal(0, 30, n), 3000 * np.sin(t / 4.0) + rng.normal(0, 300, n)], axis=1),
        sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=2),
    ]
    for x in sigs:
        x = np.clip(np.round(x), -32768, 32767).astype(np.int32)
        base, _ = O.encode_frames(O.default_config(), x, 2, 16, 44100, 4096, nthreads=8)
        for k in (1, 4, 8):
            _compare(x, 2, 16, 44100, 4096, ext_lpc_order_search=k)
            ext, _ = O.encode_frames(O.default_config(ext_lpc_order_search=k), x, 2, 16, 44100, 4096, nthreads=8)
            assert len(ext) <= len(base)
    x = np.clip(np.round(sigs[0]), -32768, 32767).astype(np.int32)
    small, _ = O.encode_frames(O.default_config(ext_lpc_order_search=4), x, 2, 16, 44100, 4096, nthreads=8)
    base, _ = O.encode_frames(O.default_config(), x, 2, 16, 44100, 4096, nthreads=8)
    assert len(small) < len(base)
    _compare(x, 2, 16, 44100, 4096, ext_lpc_order_search=8, lpc_order=24)
    _compare(x, 2, 16, 44100, 1000, ext_lpc_order_search=3, lpc_order=3, quant_precision=7)
    _compare(x[:, :1], 1, 16, 44100, 512, ext_lpc_order_search=2, lpc_order=1)
    _compare(x, 2, 16, 44100, 4096, ext_lpc_order_search=2, fixed_order_sel=0)
    _compare(x, 2, 16, 44100, 4096, ext_lpc_order_search=2, use_fixed=0, window_type=0)
    y = sigen.noisy_sine_pcm(4608 * 2 + 100, 3, 24, 96000, config_id=3)
    _compare(y, 3, 24, 96000, 4608, ext_lpc_order_search=4, lpc_order=16)
    z = sigen.noisy_sine_pcm(1024 * 3 + 9, 8, 24, 48000, config_id=5)
    _compare(z, 8, 24, 48000, 1024, ext_lpc_order_search=2)
    _compare(np.zeros((3000, 2), np.int32), 2, 16, 44100, 1024, ext_lpc_order_search=2, use_constant=0)
    _compare(crafted_huge_residual_stereo(), 2, 24, 96000, 4096, lpc_order=24, ext_lpc_order_search=3)  # a fallback frame
    # second extension: fewer quantiser bits for the order-P coefficients, alone and with the order search
    for ko, kp in ((0, 1), (0, 4), (4, 4), (5, 3)):
        _compare(x, 2, 16, 44100, 4096, ext_lpc_order_search=ko, ext_lpc_precision_search=kp)
        ext, _ = O.encode_frames(O.default_config(ext_lpc_order_search=ko, ext_lpc_precision_search=kp), x, 2, 16, 44100, 4096,
                                 nthreads=8)
        assert len(ext) <= len(base)
    _compare(x, 2, 16, 44100, 1000, ext_lpc_precision_search=4, quant_precision=3, lpc_order=5)
    _compare(y, 3, 24, 96000, 4608, ext_lpc_order_search=3, ext_lpc_precision_search=2, lpc_order=16)
    _compare(crafted_huge_residual_stereo(), 2, 24, 96000, 4096, lpc_order=24, ext_lpc_precision_search=2)
    with pytest.raises(VerifyError):
        make_config(ext_lpc_order_search=6, ext_lpc_precision_search=3).into_verified()
    with pytest.raises(VerifyError):
        make_config(ext_lpc_order_search=1, use_direct_mse=1).into_verified()
    with pytest.raises(VerifyError):
        make_config(ext_lpc_order_search=9).into_verified()


def test_large_blocks_with_256_finest_partitions():
    """block sizes 30720 .. 32512 in steps of 256 (2^8 finest Rice partitions at nearly the maximum length): the generic
    rice kernel fits them in shared memory only without its bank padding -- found by the fuzz campaign (31744), where the
    library answered with its "shared memory budget exceeded" error"""
    rng = np.random.default_rng(77)
    for block in (31744, 32512, 30720, 32000):
        x = (rng.normal(0, 2000, (block + 300, 2)).cumsum(axis=0) % 60000 - 30000).astype(np.int32)
        _compare(x, 2, 16, 44100, block)
    x4 = (rng.normal(0, 200000, (31744, 4))).astype(np.int32)
    _compare(x4, 4, 24, 44100, 31744, lpc_order=6, quant_precision=10, fixed_order_sel=0, prc_max_parameter=6, use_midside=0,
             ext_lpc_order_search=6, ext_lpc_precision_search=2)
    x8 = (rng.normal(0, 100000, (32512 + 50, 8))).astype(np.int32)
    _compare(x8, 8, 24, 48000, 32512)


def test_pipelined_host_path_matches_oracle(monkeypatch):
    """Host batches longer than 1.5 chunks go through the H2D / kernels / D2H pipeline (three rotating buffer
    sets, two compute streams).  With 5-frame chunks a 38-frame batch exercises set reuse several times."""
    monkeypatch.setenv("FB200_CHUNK_FRAMES", "5")
    vcfg = Encoder().into_verified()
    n = 4096 * 37 + 1000
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=9)
    ref, ref_sizes = O.encode_frames(O.default_config(), x, 2, 16, 44100, 4096, first_frame_number=7)
    with Context(vcfg, 2, 16, 44100, 4096) as ctx:
        for _ in range(2):  # second call reuses the sets
            got, sizes, infos = ctx.encode_interleaved(pack_pcm(x, 2), 2, n, 7, want_infos=True)
            assert list(sizes) == list(ref_sizes)
            assert got.tobytes() == ref
            assert [infos[i].frame_number for i in range(38)] == list(range(7, 45))
            assert [infos[i].frame_bytes for i in range(38)] == list(ref_sizes)
            t = ctx.timing()
            assert (t.fused_frames, t.fallback_frames) == (38, 0)  # (tail 1000 = 8 x 125)
            assert t.launches == 9 * 8  # 9 chunks; 16-bit stereo is read as PCM pairs: no ingest kernel
        # an out-of-range sample in a late chunk is still a VerifyError
        bad = x.copy()
        bad[4096 * 30 + 5, 1] = 40000
        with pytest.raises(VerifyError):
            ctx.encode_interleaved(pack_pcm(bad, 4), 4, n)
        # and the context stays usable afterwards
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, 2), 2, n, 7)
        assert got.tobytes() == ref
        # capacity error
        small = np.empty(len(ref) - 10, np.uint8)
        nf, olen = C.c_size_t(0), C.c_size_t(0)
        buf = pack_pcm(x, 2)
        rc = _ffi.lib().fb200_encode_interleaved(ctx._h, buf.ctypes.data, 2, n, 7, small.ctypes.data, len(small), None,
                                                  None, C.byref(nf), C.byref(olen))
        assert rc == _ffi.ERR_CAPACITY
    out, nf = O.decode_frames(ref, 2, 16)
    assert nf == 38 and np.array_equal(out, x)


def test_pipelined_chunk_schedule(monkeypatch):
    """even body chunks plus one short final chunk (123, 123, 122 + 32 frames here): sizes, bytes and order are intact"""
    monkeypatch.setenv("FB200_CHUNK_FRAMES", "128")
    n = 1024 * 400
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=11)
    ref, ref_sizes = O.encode_frames(O.default_config(), x, 2, 16, 44100, 1024)
    with Context(Encoder().into_verified(), 2, 16, 44100, 1024) as ctx:
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, 2), 2, n)
        assert ctx.timing().launches == 4 * 8  # 4 chunks x 8 kernels (no ingest kernel for 16-bit stereo)
    assert list(sizes) == list(ref_sizes) and got.tobytes() == ref


def test_odd_block_size_every_frame_on_the_odd_instances():
    """block 1000 = 8 partitions of 125 samples: no unit of any frame starts on a multiple of 4 samples"""
    n = 1000 * 7 + 333
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=9)
    _compare(x, 2, 16, 44100, 1000)
    with Context(make_config(block_size=1000).into_verified(), 2, 16, 44100, 1000) as ctx:
        ctx.encode_interleaved(pack_pcm(x, 2), 2, n)
        t = ctx.timing()
        assert (t.fused_frames, t.fallback_frames) == (8, 0)


def test_fused_geometry_odd_block_sizes():
    rng = np.random.default_rng(11)
    for n in (64, 66, 127, 128, 341 * 8, 3136, 98 * 32, 5000, 4100, 8192, 9216, 12345, 16383):
        x = (rng.normal(0, 300, n + 17).cumsum() % 20000 - 10000).astype(np.int32)
        _compare(x, 1, 16, 44100, n)


@pytest.mark.parametrize("seed", range(40))
def test_randomised_formats_signals_and_configs(seed):
    """the seeded fuzz of tests/test_kernel_logic_emu.py against the real kernels (both device paths)"""
    rng = np.random.default_rng(1000 + seed)
    x, channels, bps, rate, block, first, cfg = random_case(rng)
    _compare(x, channels, bps, rate, block, first_frame=first, **cfg)


@pytest.mark.parametrize("seed", range(60))
def test_fuzz_frames_at_reference_scale(seed):
    """the property of the reference's fuzz target (fuzz/fuzz_targets/frame_encode.rs:197-212) at its scale -- one frame of
    32..32767 samples, 8..24 bits, 1..8 channels, every configuration toggle random -- strengthened from "decodes
    losslessly" to "bytes equal the oracle's on every device path"; a second frame (shorter) exercises the tail"""
    rng = np.random.default_rng(9000 + seed)
    x, channels, bps, rate, block, cfg = fuzz_frame_case(rng)
    tail = x[: int(rng.integers(1, block + 1))]
    _compare(np.concatenate([x, tail]), channels, bps, rate, block, **cfg)


def test_first_frame_number_and_utf8_lengths():
    x = sigen.Sine(50, 0.3).noise(0.01, seed=1).to_vec_quantized(16, 600).reshape(300, 2)
    for first in (0, 127, 128, 2047, 65535, (1 << 21) - 1, (1 << 26), (1 << 31) - 3):
        _compare(x, 2, 16, 44100, 128, first_frame=first)


def test_errors_map_to_reference_error_kinds():
    cfg = Encoder().into_verified()
    x = np.zeros((200, 2), np.int32)
    x[77, 1] = 2048
    with Context(cfg, 2, 12, 44100, 64) as ctx:
        with pytest.raises(VerifyError):          # sample out of range (src/source.rs:262-275)
            ctx.encode_interleaved(pack_pcm(x, 2), 2, 200)
        x[77, 1] = -2048
        ctx.encode_interleaved(pack_pcm(x, 2), 2, 200)
        with pytest.raises(VerifyError):          # frame number >= 2^31 (src/coding.rs:587-591)
            ctx.encode_interleaved(pack_pcm(x, 2), 2, 200, (1 << 31) - 2)
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, 2), 2, 0)   # empty input: zero frames
        assert len(got) == 0 and len(sizes) == 0
    bad = Encoder()
    bad.subframe_coding.qlpc.lpc_order = 25
    with pytest.raises(VerifyError):
        bad.into_verified()
    with pytest.raises(VerifyError):
        Context(cfg, 9, 16, 44100, 4096)


def test_float_tier_taps():
    """north_star tier 2: autocorrelation and LPC floats within 1e-5 relative (bit-identical in practice),
    quantised coefficients exact"""
    x = sigen.noisy_sine_pcm(4096 * 2 + 999, 2, 16, 44100)
    with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
        taps, nv = ctx.analyze(pack_pcm(x, 2), 2, len(x))
    assert nv == 12
    for f in range(3):
        blk = x[f * 4096:(f + 1) * 4096]
        L, R = blk[:, 0], blk[:, 1]
        for v, sig in enumerate([L, R, (L + R) >> 1, L - R]):
            coefs, corr = O.lpc_from_autocorr(sig, 1, 0.4, 10)
            tp = taps[f * 4 + v]
            np.testing.assert_allclose(np.array(tp.autocorr[:11]), corr, rtol=1e-5)
            np.testing.assert_allclose(np.array(tp.lpc[:10]), coefs, rtol=1e-5, atol=1e-12)
            assert np.array_equal(np.array(tp.autocorr[:11]), corr)
            assert np.array_equal(np.array(tp.lpc[:10]), coefs)
            q, order, shift = O.quantize_parameters(coefs, 15)
            assert tp.qlp_order == order and tp.qlp_shift == shift and list(tp.qlp[:order]) == q.tolist()
            e5 = O.fixed_lpc_errors(sig)
            bps_v = 17 if v == 3 else 16
            assert list(tp.fixed_est_bits) == [O.estimate_entropy(e5[k], k, 16) + bps_v * k for k in range(5)]


def test_encode_fixed_size_frame_api():
    """encode_fixed_size_frame on a FrameBuf (src/coding.rs:581-606); short fills and decision records"""
    cfg = Encoder().into_verified()
    x = sigen.Sine(90, 0.5).noise(0.03, seed=4).to_vec_quantized(16, 2 * 1024).reshape(1024, 2)
    fb = FrameBuf.with_size(2, 1024)
    si = StreamInfo.new(44100, 2, 16)
    for filled in (1024, 700, 40):
        fb.fill_interleaved(x[:filled])
        frame = encode_fixed_size_frame(cfg, fb, 5, si)
        planar = np.zeros((2, 1024), np.int32)
        planar[:, :filled] = x[:filled].T
        ref, rec = O.encode_frame(O.default_config(), planar, 16, 44100, 5, n=filled)
        assert frame.bitstream == ref
        assert frame.channel_assignment == (rec["ch_assignment"] or 1)
        for c in range(2):
            assert frame.subframes[c].type == rec["subframes"][c]["type"]
            assert frame.subframes[c].bits == rec["subframes"][c]["bits"]
            if frame.subframes[c].type in (2, 3):
                assert frame.subframes[c].rice_params == rec["subframes"][c]["rice_params"]
            if frame.subframes[c].type == 3:
                assert frame.subframes[c].qlp == rec["subframes"][c]["qlp"]
    with pytest.raises(VerifyError):
        encode_fixed_size_frame(cfg, fb, 1 << 31, si)


def test_encode_with_fixed_block_size_stream():
    """encode_with_fixed_block_size (src/coding.rs:645-695): whole stream incl. STREAMINFO and MD5"""
    for ch, bps, rate, bs, n in ((2, 16, 44100, 4096, 16123), (1, 16, 44100, 4096, 102), (2, 24, 16000, 128, 1024),
                                 (3, 8, 8000, 256, 5000)):
        chans = [sigen.Sine(100 + 13 * c, 0.4).noise(0.05, seed=c).to_vec_quantized(bps, n) for c in range(ch)]
        signal = np.stack(chans, axis=1)
        src = MemSource.from_samples(signal, ch, bps, rate)
        stream = encode_with_fixed_block_size(Encoder().into_verified(), src, bs)
        ref = O.encode_stream(O.default_config(), signal, ch, bps, rate, bs)
        assert stream.write() == ref
        out, info = O.decode_stream(stream.write())
        assert np.array_equal(out, signal)
        assert bytes(info.md5) == O.md5_of_samples(signal, (bps + 7) // 8)
        assert info.min_block == info.max_block and info.total_samples == n
    # the reference's MD5 known answer (src/coding.rs:737-769)
    signal = np.full((1024, 2), 23, np.int32)
    s = encode_with_fixed_block_size(Encoder().into_verified(), MemSource.from_samples(signal, 2, 24, 16000), 128)
    assert s.write()[26:42] == bytes([0xEE, 0x78, 0x7A, 0x6E, 0x99, 0x01, 0x36, 0x79, 0xA5, 0xBB, 0x6D, 0x5C, 0x10,
                                      0xAF, 0x0B, 0x87])
    # empty stream = 42 bytes (src/component/bitrepr.rs:610-621)
    e = encode_with_fixed_block_size(Encoder().into_verified(), MemSource.from_samples([], 2, 16, 44100), 4096)
    assert len(e) == 42


def test_stream_sharded_over_devices():
    """fb200_encode_stream with a device list: frame ranges shared by the devices (chunk c on device c mod N, no
    collective), bytes placed in order; min/max frame size over all frames.  A device listed twice works once."""
    ndev = _ffi.lib().fb200_device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    n = 1024 * 28 + 300
    signal = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=21)
    src = MemSource.from_samples(signal, 2, 16, 44100)
    stream = encode_with_fixed_block_size(Encoder().into_verified(), src, 1024, devices=devices)
    ref = O.encode_stream(O.default_config(), signal, 2, 16, 44100, 1024)
    assert stream.write() == ref
    out, info = O.decode_stream(stream.write())
    assert np.array_equal(out, signal) and info.total_samples == n


@pytest.mark.parametrize("n_ctx,chunk", [(2, 7), (3, 5), (5, 64)])
def test_frame_range_sharding_over_contexts(monkeypatch, n_ctx, chunk):
    """fb200_encode_interleaved_sharded: chunk c of the batch is encoded by context c mod N with that context's own
    streams and buffer sets, and its bytes land at their final offset (the offset of a chunk is only known once the
    chunks before it are done).  Contexts on different devices when the box has them; several contexts on one
    device exercise exactly the same hand-over logic.  Bytes, sizes and order must equal the oracle's."""
    from flacenc_rs_b200.encoder import encode_interleaved_sharded
    monkeypatch.setenv("FB200_CHUNK_FRAMES", str(chunk))
    ndev = max(1, _ffi.lib().fb200_device_count())
    vcfg = Encoder().into_verified()
    n = 1024 * 61 + 333
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=31)
    ref, ref_sizes = O.encode_frames(O.default_config(), x, 2, 16, 44100, 1024, first_frame_number=3)
    ctxs = [Context(vcfg, 2, 16, 44100, 1024, device=i % ndev) for i in range(n_ctx)]
    try:
        for _ in range(2):  # the second call reuses the buffer sets
            got, sizes = encode_interleaved_sharded(ctxs, pack_pcm(x, 2), 2, n, 3)
            assert list(sizes) == list(ref_sizes)
            assert got.tobytes() == ref
        t = ctxs[0].timing()
        assert t.fused_frames == 62 and t.total_ms > 0
        # a late out-of-range sample is still a VerifyError, and the contexts stay usable
        bad = x.copy()
        bad[1024 * 50 + 5, 0] = 40000
        with pytest.raises(VerifyError):
            encode_interleaved_sharded(ctxs, pack_pcm(bad, 4), 4, n, 3)
        got, sizes = encode_interleaved_sharded(ctxs, pack_pcm(x, 4), 4, n, 3)
        assert got.tobytes() == ref
    finally:
        for c in ctxs:
            c.close()


def test_stream_api_rejects_samples_outside_bits_per_sample():
    """FrameBuf::verify_samples (/root/reference/src/source.rs:262-275): a sample that does not fit bits_per_sample is a
    VerifyError -- also one that packing to ceil(bps / 8) bytes would silently wrap (40000 at 16 bits -> -25536): the
    stream-level mirror hands the int32 samples to the library unpacked, so the device's range check sees true values."""
    from flacenc_rs_b200.encoder import encode_streams_with_fixed_block_size
    for bps, bad in ((16, 40000), (16, -32769), (8, 128), (24, 1 << 23), (12, 2048), (12, -2049)):
        x = np.zeros((300, 2), np.int32)
        x[123, 1] = bad
        with pytest.raises(VerifyError):
            encode_with_fixed_block_size(Encoder().into_verified(), MemSource.from_samples(x, 2, bps, 44100), 128)
        with pytest.raises(VerifyError):
            encode_streams_with_fixed_block_size(Encoder().into_verified(), [MemSource.from_samples(x, 2, bps, 44100)], 128)


def test_batch_of_streams():
    """fb200_encode_streams: every stream of a batch equals the oracle's stream (MD5 threads run next to the device
    work); lengths differ, one stream is shorter than a block"""
    from flacenc_rs_b200.encoder import encode_streams_with_fixed_block_size
    ndev = max(1, _ffi.lib().fb200_device_count())
    lens = [4096 * 9 + 17, 100, 4096 * 3, 4096 * 20 + 4000, 4096]
    sigs = [sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=40 + i) for i, n in enumerate(lens)]
    srcs = [MemSource.from_samples(sg, 2, 16, 44100) for sg in sigs]
    streams = encode_streams_with_fixed_block_size(Encoder().into_verified(), srcs, 4096, devices=list(range(min(ndev, 2))))
    assert len(streams) == len(lens)
    for sg, st in zip(sigs, streams):
        assert st.write() == O.encode_stream(O.default_config(), sg, 2, 16, 44100, 4096)
    _ffi.lib().fb200_pool_clear()


def test_device_resident_api_matches_host_api():
    """fb200_encode_device (HBM in, HBM out) returns the same bytes as the host-buffer call"""
    import torch
    x = sigen.noisy_sine_pcm(4096 * 37 + 123, 2, 16, 44100, config_id=2)
    pcm = pack_pcm(x, 2)
    with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
        host_bytes, host_sizes, _ = ctx.encode_interleaved(pcm, 2, len(x))
        host_bytes = host_bytes.tobytes()
        d_in = torch.from_numpy(pcm.copy()).cuda()
        d_out = torch.empty(38 * ctx.max_frame_bytes(), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        olen, sizes = ctx.encode_device(d_in.data_ptr(), 2, len(x), d_out.data_ptr(), d_out.numel())
        assert olen == len(host_bytes) and list(sizes) == list(host_sizes)
        assert d_out[:olen].cpu().numpy().tobytes() == host_bytes


@pytest.mark.parametrize("fmt", [(2, 16, 2), (2, 24, 3), (3, 16, 2)])
def test_device_resident_api_in_several_chunks(monkeypatch, fmt):
    """the device-resident (serial) path cut into chunks by its working-set bound (FB200_SERIAL_MIB; default 4 GiB, so an
    hour of CD stereo is one chunk): offsets continue across chunks on the device, bytes equal the oracle's"""
    import torch
    channels, bps, cont = fmt
    monkeypatch.setenv("FB200_SERIAL_MIB", "16")  # ~190 stereo frames of 4096 samples per chunk
    x = sigen.noisy_sine_pcm(4096 * 450 + 77, channels, bps, 44100, config_id=2)
    pcm = pack_pcm(x, cont)
    ref, ref_sizes = O.encode_frames(O.default_config(), x, channels, bps, 44100, 4096, nthreads=8)
    with Context(Encoder().into_verified(), channels, bps, 44100, 4096) as ctx:
        d_in = torch.from_numpy(pcm.copy()).cuda()
        d_out = torch.empty(451 * ctx.max_frame_bytes(), dtype=torch.uint8, device="cuda")
        olen, sizes = ctx.encode_device(d_in.data_ptr(), cont, len(x), d_out.data_ptr(), d_out.numel())
        assert ctx.timing().launches > 16  # more than two chunks' worth of launches
        assert list(sizes) == list(ref_sizes)
        assert d_out[:olen].cpu().numpy().tobytes() == ref


def test_device_resident_api_unaligned_pcm():
    """a device PCM pointer that is only 4-byte aligned: the ingest kernel's generic loads and the pack kernel's plane
    staging (the PCM-pair staging needs 16-byte alignment) give the same bytes"""
    import torch
    x = sigen.noisy_sine_pcm(4096 * 5 + 999, 2, 16, 44100, config_id=4)
    pcm = pack_pcm(x, 2)
    with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
        host_bytes, host_sizes, _ = ctx.encode_interleaved(pcm, 2, len(x))
        host_bytes = host_bytes.tobytes()
        d_buf = torch.zeros(len(pcm) + 64, dtype=torch.uint8, device="cuda")
        d_buf[4:4 + len(pcm)] = torch.from_numpy(pcm.copy()).cuda()
        d_out = torch.empty(6 * ctx.max_frame_bytes(), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        olen, sizes = ctx.encode_device(d_buf.data_ptr() + 4, 2, len(x), d_out.data_ptr(), d_out.numel())
        assert olen == len(host_bytes) and list(sizes) == list(host_sizes)
        assert d_out[:olen].cpu().numpy().tobytes() == host_bytes


def test_large_batch_properties_c2_slice():
    """size-independent checks on a large batch (2000 frames of config 2): lossless decode, both CRCs valid
    (the decoder verifies them), sizes sum to the stream, stream size equals the oracle's on a sample"""
    n = 4096 * 2000 + 3136
    x = sigen.noisy_sine_pcm(n, 2, 16, 44100, config_id=2)
    with Context(Encoder().into_verified(), 2, 16, 44100, 4096) as ctx:
        got, sizes, _ = ctx.encode_interleaved(pack_pcm(x, 2), 2, n)
    assert int(sizes.sum()) == len(got) and len(sizes) == 2001
    out, nf = O.decode_frames(got.tobytes(), 2, 16)
    assert nf == 2001 and np.array_equal(out, x)
    # frame bytes of a random sample of frames against the oracle
    offs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))]).astype(np.int64)
    rng = np.random.default_rng(0)
    for f in sorted(set(rng.integers(0, 2001, 40).tolist() + [0, 2000])):
        blk = x[f * 4096:(f + 1) * 4096]
        ref, _ = O.encode_frames(O.default_config(), blk, 2, 16, 44100, 4096, first_frame_number=f)
        assert got[offs[f]:offs[f + 1]].tobytes() == ref, f


def test_device_log2f_matches_libm_exhaustive():
    """The device build of fb_log2f (estimate_entropy's log2, /root/reference/src/coding.rs:200-227) against the host
    glibc log2f for every non-negative float bit pattern + a slice of the negative ones: bit-equal or both NaN."""
    lib = _ffi.lib()
    threads = min(32, os.cpu_count() or 1)
    step = 1 << 26
    ranges = [(f, min(step, 0x7F800000 + 4096 - f)) for f in range(0, 0x7F800000 + 4096, step)] + [(0x80000000, 1 << 22)]
    out = np.empty(step, np.uint32)
    for first, cnt in ranges:
        assert lib.fb200_debug_log2f(0, first, cnt, out.ctypes.data) == 0
        ref = O.log2f_bits(first, cnt, threads)
        got = out[:cnt]
        neq = got != ref
        if neq.any():
            both_nan = np.isnan(got.view(np.float32)) & np.isnan(ref.view(np.float32))
            bad = np.flatnonzero(neq & ~both_nan)
            assert len(bad) == 0, f"{len(bad)} mismatches, first at bits {first + int(bad[0]):#010x}"


def test_device_find_shift_matches_libm_around_powers_of_two():
    """fb_find_shift on the device == ceil(log2()) through the host libm (the oracle), incl. the inexact cases
    (/root/reference/src/lpc.rs:234-255)"""
    lib = _ffi.lib()
    vals = np.array(shift_cases(), np.float64)
    vals = np.concatenate([vals, -vals])
    for prec in (1, 4, 8, 15):
        got = np.empty(len(vals), np.int32)
        assert lib.fb200_debug_find_shift(0, vals.ctypes.data, len(vals), prec, got.ctypes.data) == 0
        want = O.find_shift_each(vals, prec)
        assert np.array_equal(got, want), np.flatnonzero(got != want)[:5]


def test_cli_encodes_wav_files(tmp_path):
    """the encode command of the reference's binary (flacenc-bin/src/main.rs:204-269) on the library: WAV in (16-bit
    stereo, 24-bit mono, 8-bit unsigned), FLAC out; the streams equal the oracle's and decode to the WAV's samples"""
    import wave
    from flacenc_rs_b200 import cli
    cases = [(2, 2, 44100, 4096 * 3 + 555), (1, 3, 48000, 5000), (2, 1, 8000, 3000)]
    paths, sigs = [], []
    for i, (ch, width, rate, n) in enumerate(cases):
        bps = width * 8
        x = sigen.noisy_sine_pcm(n, ch, bps, rate, config_id=50 + i)
        raw = pack_pcm(x, width)
        if width == 1:
            raw = (raw.view(np.int8).astype(np.int32) + 128).astype(np.uint8)  # 8-bit WAV is unsigned
        path = str(tmp_path / f"in{i}.wav")
        with wave.open(path, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(width); w.setframerate(rate)
            w.writeframes(raw.tobytes())
        paths.append(path)
        sigs.append((x, ch, bps, rate))
    for i, path in enumerate(paths):
        out = str(tmp_path / f"out{i}.flac")
        assert cli.main(["encode", "-o", out, path]) == 0
        x, ch, bps, rate = sigs[i]
        data = open(out, "rb").read()
        assert data == O.encode_stream(O.default_config(), x, ch, bps, rate, 4096)
        dec, info = O.decode_stream(data)
        assert np.array_equal(dec, x)
    # --config: the reference's TOML form (here report/experimental.config.toml's content), --dump-config writes it back
    toml = str(tmp_path / "experimental.config.toml")
    with open(toml, "w") as f:
        f.write("[subframe_coding.qlpc]\nuse_direct_mse = true\n\n[subframe_coding.qlpc.window]\ntype = \"Rectangle\"\n")
    out = str(tmp_path / "c4.flac")
    dump = str(tmp_path / "used.toml")
    assert cli.main(["encode", "-c", toml, "--dump-config", dump, "-o", out, paths[0]]) == 0
    x, ch, bps, rate = sigs[0]
    assert open(out, "rb").read() == O.encode_stream(O.default_config(use_direct_mse=1, window_type=0), x, ch, bps, rate, 4096)
    used = Encoder.from_toml(open(dump).read())
    assert used.subframe_coding.qlpc.use_direct_mse and used.subframe_coding.qlpc.window.type == "Rectangle"
    # two files of one format as a batch
    assert cli.main(["encode", "-o", str(tmp_path / "b"), paths[0], paths[0]]) == 0
    assert open(str(tmp_path / "b0.flac"), "rb").read() == open(str(tmp_path / "b1.flac"), "rb").read() == open(str(tmp_path / "out0.flac"), "rb").read()
